// qk_guided.cu -- SURVEY 8f-4: the reference's fitted-linear evaluation (evaluation.py:130-212) in batch,
// and playouts under MCTSEngine._select_rollout_move's eval-guided epsilon-greedy policy (mcts.py:345-383).
// One position / one playout per thread.  A guided ply scores every child of the position (up to 64 evaluations);
// the evaluation looks at all 12 lines at once in packed 4-bit fields that the playout carries along
// (qk_eval_core.cuh: EvalPos), so a child costs about 100 integer instructions plus 11 float64 operations.
#include "qk_common.cuh"
#include "qk_eval_core.cuh"

namespace qk {

static constexpr unsigned kBlock = 128;

struct EvalWeights { double w[6]; };

__global__ void __launch_bounds__(kBlock) k_eval_features(const uint4 *__restrict__ states, const uint8_t *__restrict__ player,
                                                           size_t n, double *__restrict__ features)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        const uint4 v = __ldg(states + i);
        const State4 s = { v.x, v.y, v.z, v.w };
        int f[6];
        eval_features(s, player[i] & 1, f);
#pragma unroll
        for (int k = 0; k < 6; ++k) features[6 * i + k] = (double)f[k];
    }
}

// the greedy line of every root (qk_eval_core.cuh: GuidedLine), one thread per root
__global__ void __launch_bounds__(kBlock) k_guided_lines(const uint4 *__restrict__ roots, uint64_t n_roots, const EvalWeights weights,
                                                          GuidedLine *__restrict__ lines)
{
    const uint64_t i = blockIdx.x * (uint64_t)kBlock + threadIdx.x;
    if (i >= n_roots) return;
    const uint4 v = __ldg(roots + i);
    const State4 s = { v.x, v.y, v.z, v.w };
    lines[i] = guided_line(s, weights.w);
}

__global__ void __launch_bounds__(kBlock) k_playout_guided(const uint4 *__restrict__ roots, uint64_t n_roots, uint32_t rollouts,
                                                            uint64_t seed, uint64_t first_rollout, uint32_t root_index_base,
                                                            int max_plies, double epsilon, const EvalWeights weights,
                                                            const GuidedLine *__restrict__ lines,
                                                            uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws,
                                                            int8_t *per_playout, uint8_t *per_plies)
{
    const uint64_t total = n_roots * rollouts;
    for (uint64_t id = blockIdx.x * (uint64_t)kBlock + threadIdx.x; id < total; id += (uint64_t)gridDim.x * kBlock) {
        const uint32_t root = (uint32_t)(id / rollouts), j = (uint32_t)(id % rollouts);
        const uint4 v = __ldg(roots + root);
        const State4 s = { v.x, v.y, v.z, v.w };
        int plies;
        const int res = guided_playout_one(s, seed, first_rollout + j, root_index_base + root, max_plies, epsilon, weights.w, &plies,
                                           lines ? lines + root : nullptr);
        // tallies: the lanes of the warp that played for the same root (consecutive ids: almost always all of them)
        // are summed first, one lane per root issues the atomics
        const unsigned peers = __match_any_sync(__activemask(), root);
        const unsigned p0 = __ballot_sync(peers, res > 0), p1 = __ballot_sync(peers, res < 0);
        if ((unsigned)(__ffs((int)peers) - 1) == (threadIdx.x & 31u)) {
            const uint32_t n0 = (uint32_t)__popc(p0 & peers), n1 = (uint32_t)__popc(p1 & peers), nd = (uint32_t)__popc(peers) - n0 - n1;
            if (wins_p0 && n0) atomicAdd(wins_p0 + root, n0);
            if (wins_p1 && n1) atomicAdd(wins_p1 + root, n1);
            if (draws && nd) atomicAdd(draws + root, nd);
        }
        if (per_playout) per_playout[id] = (int8_t)res;
        if (per_plies) per_plies[id] = (uint8_t)plies;
    }
}

int launch_eval_features(const uint4 *states, const uint8_t *player, size_t n, double *features, Ctx *c, cudaStream_t s)
{
    k_eval_features<<<grid_for(n, kBlock, c->sms, 16), kBlock, 0, s>>>(states, player, n, features);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

size_t guided_scratch_bytes(size_t n_roots) { return n_roots * sizeof(GuidedLine); }

int launch_playouts_guided(const uint4 *roots, size_t n_roots, uint32_t rollouts, uint64_t seed, uint64_t first_rollout,
                           uint32_t root_index_base, int max_plies, double epsilon, const double *weights, uint32_t *wins_p0,
                           uint32_t *wins_p1, uint32_t *draws, int8_t *per_playout, uint8_t *per_plies, void *line_scratch, Ctx *c,
                           cudaStream_t s)
{
    EvalWeights w;
    for (int i = 0; i < 6; ++i) w.w[i] = weights[i];
    const size_t total = n_roots * (size_t)rollouts;
    // the greedy line pays off as soon as a root is played a few times (it costs about one non-exploring playout);
    // with epsilon = 1 no playout ever follows it
    GuidedLine *lines = (line_scratch && rollouts >= 4 && epsilon < 1.0) ? (GuidedLine *)line_scratch : nullptr;
    if (lines) {
        k_guided_lines<<<(unsigned)((n_roots + kBlock - 1) / kBlock), kBlock, 0, s>>>(roots, n_roots, w, lines);
        QK_CUDA(cudaGetLastError());
    }
    k_playout_guided<<<grid_for(total, kBlock, c->sms, 16), kBlock, 0, s>>>(roots, n_roots, rollouts, seed, first_rollout,
                                                                            root_index_base, max_plies, epsilon, w, lines, wins_p0,
                                                                            wins_p1, draws, per_playout, per_plies);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

}  // namespace qk
