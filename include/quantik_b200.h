/*
 * quantik_b200.h -- C ABI of libquantik_b200.so: the B200 (sm_100a) implementation
 * of quantik-core's batched-simulation hot path.
 *
 * The reference (mberlanda/quantik-core-py 1.1.0) is pure Python and has no FFI
 * layer; its seam is the set of module-level functions the engines import by name
 * (mcts.py:15-19, beam_search.py:20-22, minimax.py:49-51, game_stats.py:7-22) plus
 * the batch wire format of storage/compact_state.py:125-172.  Every entry point
 * below cites the reference function it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *  - A state is 8 little-endian uint16 words [P0.A P0.B P0.C P0.D P1.A P1.B P1.C
 *    P1.D], bit i = square i, row-major 4x4 (commons.py:10-13;
 *    memory/bitboard_compact.py:36-37 `<8H`).  Batches are N x 8 words, 16-byte
 *    aligned (storage/compact_state.py:125-147).
 *  - An action is shape*16 + square (action-index.v1, docs/API_PORTABILITY.md:21;
 *    api_portability_report.py:75-76); a legal-move set is the 64-bit mask of
 *    actions in that numbering, which is also the reference's move order
 *    (move.py:247,262-266).
 *  - A transform is d4 | colour_swap << 3 | perm[0] << 4 | perm[1] << 6 |
 *    perm[2] << 8 | perm[3] << 10, fields as in SymmetryTransform
 *    (symmetry.py:45-62).
 *  - Every pointer may be HOST memory (numpy) or DEVICE memory (torch data_ptr);
 *    the library asks cudaPointerGetAttributes.  Host buffers are staged through
 *    the device's stream-ordered pool and the call returns when results are in
 *    place; with device buffers only, the call is asynchronous on `stream`.
 *  - `dev` is the CUDA device ordinal, `stream` a cudaStream_t (NULL = default
 *    stream).  Return value 0 = success, negative = QK_E_*; qk_last_error() has
 *    the text (thread local).  There is no CPU fallback: without a usable sm_100
 *    device every compute entry point fails with QK_E_CUDA / QK_E_NO_DEVICE.
 *  - No global RNG state: a playout's moves are a pure function of
 *    (seed, root index, rollout index, ply) -- see qk_playouts.
 */
#ifndef QUANTIK_B200_H
#define QUANTIK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QK_ABI_VERSION 1

#if defined(__GNUC__)
#define QK_API __attribute__((visibility("default")))
#else
#define QK_API
#endif

enum {
    QK_OK = 0,
    QK_E_BAD_ARG = -1,     /* null pointer, n out of range, depth out of range ...      */
    QK_E_CUDA = -2,        /* a CUDA runtime call failed; text in qk_last_error()        */
    QK_E_NO_DEVICE = -3,   /* no CUDA device / not an sm_100 part / qk_init not called   */
    QK_E_NOMEM = -4
};

/* WinStatus, game_utils.py:32-37 */
enum { QK_NO_WIN = 0, QK_PLAYER_0_WINS = 1, QK_PLAYER_1_WINS = 2 };
/* ValidationResult, state_validator.py:16-27 */
enum { QK_V_OK = 0, QK_V_TURN_BALANCE_INVALID = 1, QK_V_SHAPE_COUNT_EXCEEDED = 2, QK_V_NOT_PLAYER_TURN = 3,
       QK_V_ILLEGAL_PLACEMENT = 4, QK_V_PIECE_OVERLAP = 5 };

#define QK_MAX_DEPTH 16

/* ---- lifetime ------------------------------------------------------------------ */
QK_API int qk_abi_version(void);
/* Creates the per-device context (stream-ordered memory pool kept warm, SM count). */
QK_API int qk_init(int dev);
QK_API int qk_shutdown(int dev);
QK_API const char *qk_last_error(void);
/* Tuning / testing switches, process-wide; every option defaults to 0 = the library's own choice, and
 * the library never reads the environment.  They select between kernels that compute the same result
 * (tests/ uses them to put every variant through the parity checks, scripts/ for A/B timing):
 *   "perft_kernel"        0 auto | 1 k_perft_tile | 2 k_perft_lane | 3 k_perft_dfs (warp per item)
 *   "perft_min_items"     > 0: start the depth-first kernel as soon as the frontier has this many items
 *   "perft_tile_generic"  1: k_perft_tile without the one-side / 32-bit-multiplicity fast form
 *   "perft_own_lanes"     k_perft_tile, lanes evaluating children of their own instead of dealing them: 0 auto (wide
 *                         trees only) | 33 never | n = 1..32 always, from n giving lanes of a warp on, loops as long
 *                         as the shortest list | 64 + n: ... as the longest list | 128 + n: ... as the mean length
 *   "playout_variant"     0 auto | 2 the general kernel for every case | 3 / 4 the predicated tallies-only kernels at
 *                         5 / 6 resident CTAs per SM instead of 4 | 1 / 6 the older single-root kernels
 *   "dedup_mode"          0 auto | 1 one-level table insert | 2 partitioned (two-level) insert
 *   "dedup_part_mb"       > 0: size in MiB of the table one part of the two-level insert is merged in (0 = 112)
 *   "dedup_streams"       1..4: part tables in use at a time, each on its own stream (0 = 3)
 *   "table_slots_x10"     >= 11: slots of the one-level merge table per expected class, times ten (0 = 15)
 *   "guided_scoring"      0 children dealt out over the warp | 1 every lane scores its own children
 *   "call_stage"          1: host buffers and scratch of small calls go through the memory pool like large ones, not
 *                         through the per-device staging block
 *   "grid_waves"          > 0: waves of CTAs a grid-stride kernel is launched with (0 = 16)
 *   "playout_grab"        > 0: playouts a lane of the single-root playout kernel takes from the work queue at a time  */
QK_API int qk_set_option(const char *name, long long value);
/* SM count and clock (kHz) of an initialised device; either pointer may be NULL. */
QK_API int qk_device_info(int dev, int *sm_count, int *sm_clock_khz);

/* ---- rules primitives, one state per thread --------------------------------------
 * qk_movegen_masks  <- move.generate_legal_moves / generate_legal_moves_list
 *                      (move.py:199-291) + state_validator._validate_game_state_single_pass
 *                      (state_validator.py:126-168).  to_move = side to move, status =
 *                      ValidationResult; an invalid state yields mask 0, to_move 0
 *                      (move.py:223-225).  Any output pointer may be NULL.              */
QK_API int qk_movegen_masks(const uint16_t *states, size_t n, uint64_t *action_mask, uint8_t *to_move,
                     uint8_t *status, int dev, void *stream);
/* qk_apply_moves    <- move.apply_move (move.py:135-166); no legality check, like the
 *                      reference.  action = shape*16 + square places a piece of the side to
 *                      move; action | 0x80 | player << 6 places it for an explicit player
 *                      (Move.player, move.py:24-39).                                     */
QK_API int qk_apply_moves(const uint16_t *states, const uint8_t *action, size_t n, uint16_t *out,
                   int dev, void *stream);
/* qk_win_status     <- game_utils.check_game_winner / has_winning_line
 *                      (game_utils.py:123-189): WinStatus per state.                     */
QK_API int qk_win_status(const uint16_t *states, size_t n, uint8_t *win_status, int dev, void *stream);

/* ---- symmetry ---------------------------------------------------------------------
 * qk_canonical      <- SymmetryHandler.find_canonical_form / get_canonical_payload
 *                      (symmetry.py:289-354,398-410), State.canonical_payload/key
 *                      (core.py:162-180).  payload = the 16-byte canonical `<8H`; the
 *                      18-byte key is 01 02 + payload.  color_swap = 0 is the reference
 *                      default (192 images), 1 the optional 384-image mode.  transform
 *                      (may be NULL) receives the SymmetryTransform the reference returns. */
QK_API int qk_canonical(const uint16_t *states, size_t n, int color_swap, uint16_t *payload,
                 uint16_t *transform, int dev, void *stream);
/* qk_orbit_size     <- SymmetryHandler.count_orbit_size (symmetry.py:356-396),
 *                      State.symmetry_count (core.py:182-192).                          */
QK_API int qk_orbit_size(const uint16_t *states, size_t n, uint8_t *orbit, int dev, void *stream);
/* qk_apply_symmetry <- SymmetryHandler.apply_symmetry (symmetry.py:252-287).             */
QK_API int qk_apply_symmetry(const uint16_t *states, const uint16_t *transform, size_t n, uint16_t *out,
                      int dev, void *stream);
/* qk_apply_symmetry_to_moves <- SymmetryHandler.apply_symmetry_to_move (symmetry.py:425-464), batched: move i =
 *                      (players[i], actions[i] = shape * 16 + square) under transform[i].  inverse != 0 applies
 *                      SymmetryTransform.inverse() (symmetry.py:66-85) of each code instead: with the transform
 *                      qk_canonical reports, that maps a move chosen on the canonical position (an opening-book
 *                      answer) back to the position that was looked up.                                          */
QK_API int qk_apply_symmetry_to_moves(const uint8_t *players, const uint8_t *actions, const uint16_t *transform, size_t n,
                               int inverse, uint8_t *out_players, uint8_t *out_actions, int dev, void *stream);
/* qk_canon_dedup    <- the canonical merge of game_stats.py:487-505 /
 *                      beam_search.py:526-546 / generate_opening_book.py:458-496:
 *                      canonicalise n states, merge equal payloads, sum multiplicities
 *                      (mult == NULL: 1 each).  uniq (capacity `cap` states) receives the
 *                      distinct payloads sorted in byte order, uniq_mult their summed
 *                      multiplicities, *n_uniq the count (host pointer).  Synchronous.
 *                      color_swap is a flag word here: bit 0 = the 384-image mode, bit 1
 *                      (QK_DEDUP_UNSORTED) = leave the output in table order (skips the sort;
 *                      uniq must then hold cap + 1 states, uniq_mult cap + 1 counts).      */
#define QK_DEDUP_UNSORTED 2
QK_API int qk_canon_dedup(const uint16_t *states, const uint64_t *mult, size_t n, int color_swap,
                   uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq,
                   int dev, void *stream);

/* qk_bfs_level      <- one depth of game_stats.SymmetryTable (_process_depth / _process_parent_state /
 *                      _process_move, game_stats.py:359-485) and of the opening-book BFS
 *                      (examples/generate_opening_book.py:185-280): for n non-won parent states
 *                      with multiplicities, generate every child, drop the ones that complete a
 *                      line, canonicalise the rest and merge them with summed multiplicity.
 *                      uniq / uniq_mult (capacity cap + 1) receive the next canonical frontier
 *                      (sorted in byte order when sort_output != 0, else in table order), *n_uniq
 *                      its size; stats (host, 5 values) = total moves,
 *                      line wins by P0, by P1, parents where P0 / P1 is to move and blocked, all
 *                      weighted.  QK_E_NOMEM when the frontier exceeds cap.  Synchronous.         */
QK_API int qk_bfs_level(const uint16_t *states, const uint64_t *mult, size_t n, int color_swap, int sort_output,
                        uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq,
                        uint64_t *stats, int dev, void *stream);

/* qk_book_level     <- one depth of the opening-book BFS (examples/generate_opening_book.py:185-280
 *                      _expand_chunk, :458-496 dedup) counted like opening_book_summary.build_summary
 *                      (opening_book_summary.py:36-93).  states = the canonical classes of one depth,
 *                      finished ones included.  stats (host, 2 values) = terminal classes (line complete,
 *                      or mover without a move: not expanded) and edges (distinct parent-class /
 *                      child-class pairs).  uniq (capacity cap + 1) receives the classes of the next
 *                      depth in table order, uniq_indegree their number of distinct parents, *n_uniq the
 *                      count.  QK_E_NOMEM when the next depth exceeds cap.  Synchronous.               */
QK_API int qk_book_level(const uint16_t *states, size_t n, int color_swap, uint16_t *uniq,
                         uint64_t *uniq_indegree, size_t cap, size_t *n_uniq, uint64_t *stats,
                         int dev, void *stream);

/* ---- the exchange step at N GPUs: merge fused with its all-to-all over peer memory ------------------
 * The reference merges canonical states in ONE parent process (game_stats.py:487-505,
 * examples/generate_opening_book.py:458-496).  At N GPUs equal payloads found by different ranks have to
 * meet: every payload has an owner rank (a hash of the payload), and each rank owns an INBOX that its peers
 * write into directly over NVLink -- the emit pass of the local merge stores each distinct class into the
 * owner's inbox, so no separate collective moves the data.
 *   inbox[r]      device pointer to rank r's inbox, mapped into THIS process (peer access, CUDA IPC, or
 *                 torch symmetric memory; with world == 1 simply a local buffer) -- host array of `world`
 *                 pointers.  An inbox is qk_inbox_bytes(world, slots_per_src) bytes: a header of per-source
 *                 counts and one region of slots_per_src (payload, multiplicity) entries per source rank.
 *   protocol      every rank: qk_*_scatter (returns after its peer stores are complete)  ->  a barrier among
 *                 the ranks (the caller's; e.g. an NCCL all-reduce)  ->  qk_inbox_merge on the own inbox.
 *                 A rank may scatter into an inbox again only after its owner merged it (use two inboxes
 *                 alternately when levels follow each other without a second barrier).
 * qk_canon_dedup_scatter : qk_canon_dedup's canonicalise + local merge, then scatter.  flags bit 0 = 384-image
 *                 mode.  sent (host, `world` values, may be NULL) = entries written per destination.
 * qk_bfs_level_scatter   : qk_bfs_level's expand + canonicalise + local merge (table sized for local_cap
 *                 distinct children), then scatter.  stats as qk_bfs_level.
 * qk_inbox_merge : merge the regions of the own inbox (payloads are canonical already) into uniq / uniq_mult
 *                 (capacity cap + 1; table order unless sort_output), *n_uniq = count.
 * QK_E_NOMEM: a region, the local table, or cap was too small (nothing useful was delivered: repeat the level
 * with larger sizes on EVERY rank).  world <= QK_MAX_WORLD.  All three are synchronous.                      */
#define QK_MAX_WORLD 16
QK_API int qk_inbox_bytes(int world, size_t slots_per_src, size_t *bytes);
/* qk_payload_owner : the owner rank (0..world-1) of each canonical payload -- the partition function of the
 *                 scatter kernels, for callers that route classes themselves (quantik_b200.dist's NCCL
 *                 all-to-all path uses the same function, so both exchanges agree rank by rank).            */
QK_API int qk_payload_owner(const uint16_t *payloads, size_t n, int world, uint8_t *owner, int dev, void *stream);
QK_API int qk_canon_dedup_scatter(const uint16_t *states, const uint64_t *mult, size_t n, int flags,
                                  void *const *inbox, int world, int rank, size_t slots_per_src,
                                  uint64_t *sent, int dev, void *stream);
QK_API int qk_bfs_level_scatter(const uint16_t *states, const uint64_t *mult, size_t n, int flags, size_t local_cap,
                                void *const *inbox, int world, int rank, size_t slots_per_src,
                                uint64_t *stats, uint64_t *sent, int dev, void *stream);
QK_API int qk_inbox_merge(const void *my_inbox, int world, size_t slots_per_src, int sort_output,
                          uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq, int dev, void *stream);

/* ---- random playouts ---------------------------------------------------------------
 * qk_playouts       <- BeamSearchEngine._rollout/_default_evaluate (beam_search.py:614-645)
 *                      when max_plies < 0 (play to a true terminal; result +-1), and
 *                      MCTSEngine._simulate (mcts.py:385-437, uniform branch of
 *                      _select_rollout_move :345-364) when max_plies >= 0 (stop after
 *                      max_plies moves, result 0 = the reference's 0.0).
 * For every root r < n_roots, `rollouts_per_root` playouts are run; rollout j uses the
 * Philox4x32-10 counter stream  key = seed, counter = (j + first_rollout (64 bit),
 * ply / 4, root_index_base + r), word ply % 4, move = mulhi32(word, n_legal)-th legal
 * action.  Outputs (any may be NULL): wins_p0 / wins_p1 / draws [n_roots] tallies;
 * per_playout [n_roots * rollouts] int8 outcome from player 0's view (+1, -1, 0);
 * per_plies [n_roots * rollouts] moves played.                                          */
QK_API int qk_playouts(const uint16_t *roots, size_t n_roots, uint32_t rollouts_per_root, uint64_t seed,
                int max_plies, uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws,
                int8_t *per_playout, int dev, void *stream);
QK_API int qk_playouts_ex(const uint16_t *roots, size_t n_roots, uint32_t rollouts_per_root, uint64_t seed,
                   uint64_t first_rollout, uint32_t root_index_base, int max_plies,
                   uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws,
                   int8_t *per_playout, uint8_t *per_plies, int dev, void *stream);
/* qk_eval_features  <- evaluation.features (evaluation.py:130-194): the six handcrafted features
 *                      (threat_own, threat_opp, threat_shared, mobility_diff, build_two, build_one) of
 *                      VALID positions from player[i]'s perspective, as float64 [n][6]; evaluate() is
 *                      weights @ features.
 * qk_playouts_guided<- MCTSEngine._simulate with the eval-guided policy of _select_rollout_move
 *                      (mcts.py:345-383): with probability epsilon a uniform move, else the first move
 *                      that wins or leaves the opponent without a reply, else the move whose child
 *                      scores highest under evaluate(child, mover, weights) (first maximum).  weights =
 *                      6 doubles in HOST memory (EvalConfig.weights).  Stream: Philox4x32-10 block
 *                      (rollout, 0x80000000 | ply/2, root index), word 2*(ply%2) = epsilon draw (explore
 *                      iff word * 2^-32 < epsilon), word 2*(ply%2)+1 = uniform pick.  Outputs as in
 *                      qk_playouts_ex.                                                                */
QK_API int qk_eval_features(const uint16_t *states, const uint8_t *player, size_t n, double *features,
                            int dev, void *stream);
QK_API int qk_playouts_guided(const uint16_t *roots, size_t n_roots, uint32_t rollouts_per_root, uint64_t seed,
                              uint64_t first_rollout, uint32_t root_index_base, int max_plies, double epsilon,
                              const double *weights, uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws,
                              int8_t *per_playout, uint8_t *per_plies, int dev, void *stream);
/* qk_solve          <- MinimaxEngine.solve (minimax.py:182-204; _negamax :330-453): the exact game value of
 *                      each position with mate distance.  outcome = +1 / -1: the side to move wins /
 *                      loses under optimal play; plies = number of plies until the game ends (a line
 *                      is completed, or a mover has no move) when the winner hurries and the loser
 *                      delays -- the reference's score is outcome * (EvalConfig.win - plies);
 *                      best_action = the lowest action index among the optimal moves (255 when the
 *                      position is already finished).  Exhaustive: meant for the endgame hand-off of
 *                      HybridPlayer (hybrid.py:80-122), i.e. positions with at most ~8 empty squares. */
QK_API int qk_solve(const uint16_t *states, size_t n, int8_t *outcome, uint8_t *plies, uint8_t *best_action,
                    int dev, void *stream);
/* qk_random_roots   <- benchmarks/dataset.py:38-51 (random non-terminal mid-game
 *                      positions by rejection): root i = empty board + (min_plies +
 *                      i % span) uniform-random plies, stream key seed ^ "ROOTS",
 *                      counter (i, attempt, ply/4, 0).  Writes n states for
 *                      i = first .. first+n-1.                                            */
QK_API int qk_random_roots(uint64_t seed, uint64_t first, size_t n, int min_plies, int span,
                    uint16_t *out_states, int dev, void *stream);

/* ---- exhaustive enumeration ---------------------------------------------------------
 * qk_perft          <- the counting rules of game_stats.SymmetryTable
 *                      (_process_parent_state / _process_move / _handle_no_legal_moves,
 *                      game_stats.py:384-485) as a raw tree walk below each root, weighted
 *                      by mult (NULL = 1).  Arrays have depth+1 entries, index = plies
 *                      below the root:
 *   nodes[d]             move sequences of length d ("Total Legal Moves" of depth d)
 *   wins_p0/p1[d]        sequences whose d-th move completed a line, by mover
 *   blocked_pX_loses[d]  non-won positions at depth d < depth where X must move and
 *                        cannot (the reference books them at depth d+1 as a win of the
 *                        other player); entry [depth] is 0.
 * Output pointers must be HOST memory; the call is synchronous.                          */
QK_API int qk_perft(const uint16_t *roots, const uint64_t *mult, size_t n_roots, int depth,
             uint64_t *nodes, uint64_t *wins_p0, uint64_t *wins_p1,
             uint64_t *blocked_p0_loses, uint64_t *blocked_p1_loses, int dev, void *stream);

/* qk_perft_probe    <- the cost probe behind the static partitioning of the enumeration
 *                      ("LPT/greedy by probe cost"; the reference chunks its frontier by
 *                      count, examples/generate_opening_book.py:294-332): cost[i] = number of move
 *                      sequences of length 2 below root i (0 for finished / invalid roots), one
 *                      thread per root.  quantik_b200.dist deals roots to ranks by it.             */
QK_API int qk_perft_probe(const uint16_t *roots, size_t n_roots, uint32_t *cost, int dev, void *stream);

/* ---- measurement aid ------------------------------------------------------------------
 * Runs a dependency-free integer micro-kernel and reports warp-instructions x 32 per
 * second for (a) a pure LOP3/IADD3 stream and (b) a LOP3 + IMAD mix: the measured
 * integer-issue roofline denominators used by bench.py.                                   */
QK_API int qk_int_issue_peak(int dev, double *alu_lane_ops_per_s, double *mixed_lane_ops_per_s);
/* Per-instruction-class issue rates (8 doubles, lane-ops/s): LOP3, IMAD, IMAD.HI, IMAD.WIDE, POPC,
 * LOP3+IMAD 1:1, LOP3+IMAD.HI 1:1, SHF+LOP3 1:1.  The machine model behind DESIGN.md section 5. */
QK_API int qk_pipe_rates(int dev, double *lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* QUANTIK_B200_H */
