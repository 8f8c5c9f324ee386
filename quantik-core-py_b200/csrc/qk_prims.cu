// qk_prims.cu -- batched rules primitives, one 16-byte state per thread.
// Loads and stores are single coalesced 128-bit accesses (a warp reads 512
// contiguous bytes); everything between them stays in registers (qk_bits.cuh).
#include "qk_common.cuh"
#include "qk_bits.cuh"

namespace qk {

static constexpr unsigned kBlock = 256;

__device__ __forceinline__ State4 load_state(const uint4 *p, size_t i)
{
    uint4 v = __ldg(p + i);
    State4 s = { v.x, v.y, v.z, v.w };
    return s;
}
__device__ __forceinline__ void store_state(uint4 *p, size_t i, const State4 &s)
{
    p[i] = make_uint4(s.x, s.y, s.z, s.w);
}

// move.generate_legal_moves (move.py:199-268) + state_validator (:126-168)
__global__ void __launch_bounds__(kBlock) k_movegen(const uint4 *__restrict__ states, size_t n,
                                                     uint64_t *__restrict__ mask, uint8_t *__restrict__ to_move,
                                                     uint8_t *__restrict__ status)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        int player, st;
        uint64_t m = movegen(load_state(states, i), &player, &st);
        if (mask) mask[i] = m;
        if (to_move) to_move[i] = (uint8_t)player;
        if (status) status[i] = (uint8_t)st;
    }
}

// move.apply_move (move.py:135-166); action = shape*16+pos [| 0x80 | player << 6]
__global__ void __launch_bounds__(kBlock) k_apply(const uint4 *__restrict__ states, const uint8_t *__restrict__ action,
                                                   size_t n, uint4 *__restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        State4 s = load_state(states, i);
        const uint32_t a = action[i];
        int player = 0;
        if (a & 0x80u) player = (int)((a >> 6) & 1u);          // explicit Move.player (move.py:157)
        else if (validate(s, &player) != 0) player = 0;        // side to move
        store_state(out, i, apply_action(s, player, (int)(a & 63u)));
    }
}

// game_utils.check_game_winner (game_utils.py:161-189)
__global__ void __launch_bounds__(kBlock) k_win(const uint4 *__restrict__ states, size_t n, uint8_t *__restrict__ win)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock)
        win[i] = (uint8_t)win_status(load_state(states, i));
}

// SymmetryHandler.find_canonical_form (symmetry.py:289-354): payload only
template <bool COLOR_SWAP>
__global__ void __launch_bounds__(kBlock) k_canonical(const uint4 *__restrict__ states, size_t n, uint4 *__restrict__ payload)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock)
        store_state(payload, i, canonical<COLOR_SWAP>(load_state(states, i)));
}

// ... and with the SymmetryTransform the reference reports (slow path)
__global__ void __launch_bounds__(kBlock) k_canonical_tr(const uint4 *__restrict__ states, size_t n, int color_swap,
                                                          uint4 *__restrict__ payload, uint16_t *__restrict__ transform)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        uint32_t t;
        State4 c = canonical_with_transform(load_state(states, i), color_swap != 0, &t);
        store_state(payload, i, c);
        transform[i] = (uint16_t)t;
    }
}

// SymmetryHandler.count_orbit_size (symmetry.py:356-396)
__global__ void __launch_bounds__(kBlock) k_orbit(const uint4 *__restrict__ states, size_t n, uint8_t *__restrict__ orbit)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock)
        orbit[i] = (uint8_t)orbit_size(load_state(states, i));
}

// SymmetryHandler.apply_symmetry (symmetry.py:252-287)
__global__ void __launch_bounds__(kBlock) k_apply_symmetry(const uint4 *__restrict__ states,
                                                            const uint16_t *__restrict__ transform, size_t n,
                                                            uint4 *__restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock)
        store_state(out, i, apply_symmetry(load_state(states, i), transform[i]));
}

// SymmetryHandler.apply_symmetry_to_move (symmetry.py:425-464), one move per thread: square through the D4 map
// (the board map applied to a one-square board), colour toggled by color_swap, shape relabelled to the index of
// the old shape in shape_perm.  inverse != 0 applies SymmetryTransform.inverse() of the code (symmetry.py:66-85):
// the map that takes a move found on the canonical position back to the position it was asked for.
__global__ void __launch_bounds__(kBlock) k_apply_symmetry_to_moves(const uint8_t *__restrict__ player, const uint8_t *__restrict__ action,
                                                                     const uint16_t *__restrict__ transform, size_t n, int inverse,
                                                                     uint8_t *__restrict__ out_player, uint8_t *__restrict__ out_action)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        const uint32_t t = transform[i], a = action[i] & 63u, shape = a >> 4, pos = a & 15u;
        int d4 = (int)(t & 7u);
        if (inverse && (d4 == 1 || d4 == 3)) d4 ^= 2;                       // rot90 <-> rot270, the others are involutions
        const uint32_t moved = d4_apply(1u << pos, d4) & 0xFFFFu;
        uint32_t new_shape = 0;
        if (inverse) new_shape = (t >> (4 + 2 * shape)) & 3u;              // inverse permutation: shape_inv[perm[i]] = i
        else {
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
                if (((t >> (4 + 2 * k)) & 3u) == shape) new_shape = k;      // shape_perm.index(shape)
        }
        out_player[i] = (uint8_t)((player[i] & 1u) ^ ((t >> 3) & 1u));
        out_action[i] = (uint8_t)((new_shape << 4) | (uint32_t)(__ffs((int)moved) - 1));
    }
}

// ------------------------------------------------------------------ launchers
// grids: enough CTAs for 8 resident blocks of 256 threads on every SM, then grid-stride
int launch_movegen(const uint4 *states, size_t n, uint64_t *mask, uint8_t *to_move, uint8_t *status, Ctx *c, cudaStream_t s)
{
    k_movegen<<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, n, mask, to_move, status);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}
int launch_apply(const uint4 *states, const uint8_t *action, size_t n, uint4 *out, Ctx *c, cudaStream_t s)
{
    k_apply<<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, action, n, out);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}
int launch_win(const uint4 *states, size_t n, uint8_t *win, Ctx *c, cudaStream_t s)
{
    k_win<<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, n, win);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}
int launch_canonical(const uint4 *states, size_t n, int color_swap, uint4 *payload, uint16_t *transform, Ctx *c, cudaStream_t s)
{
    unsigned grid = grid_for(n, kBlock, c->sms, 8);
    if (transform) k_canonical_tr<<<grid, kBlock, 0, s>>>(states, n, color_swap, payload, transform);
    else if (color_swap) k_canonical<true><<<grid, kBlock, 0, s>>>(states, n, payload);
    else k_canonical<false><<<grid, kBlock, 0, s>>>(states, n, payload);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}
int launch_orbit(const uint4 *states, size_t n, uint8_t *orbit, Ctx *c, cudaStream_t s)
{
    k_orbit<<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, n, orbit);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}
int launch_apply_symmetry(const uint4 *states, const uint16_t *transform, size_t n, uint4 *out, Ctx *c, cudaStream_t s)
{
    k_apply_symmetry<<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, transform, n, out);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

int launch_apply_symmetry_to_moves(const uint8_t *player, const uint8_t *action, const uint16_t *transform, size_t n, int inverse,
                                   uint8_t *out_player, uint8_t *out_action, Ctx *c, cudaStream_t s)
{
    k_apply_symmetry_to_moves<<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(player, action, transform, n, inverse, out_player, out_action);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

}  // namespace qk
