"""quantik_b200 -- B200 (sm_100a) implementation of quantik-core's batched-simulation hot path.

Public surface:
  * the reference's single-state names (``generate_legal_moves``, ``apply_move``,
    ``check_game_winner``, ``has_winning_line``, ``State``, ``Move``, ``SymmetryHandler`` ...)
    re-implemented on top of hand-written CUDA kernels (:mod:`quantik_b200.api`);
  * batched forms over N x 8 uint16 buffers, host or device (:mod:`quantik_b200.batch`);
  * engine plug-ins (:mod:`quantik_b200.engines`).
The CUDA library is loaded lazily on the first compute call and there is no CPU fallback.
"""
from .api import (  # noqa: F401
    FLAG_CANON, VERSION, Bitboard, Move, MoveValidationResult, State, SymmetryHandler, SymmetryTransform,
    ValidationError, ValidationResult, WinStatus, apply_move, bb_from_qfen, bb_to_qfen, check_game_winner,
    generate_legal_moves, generate_legal_moves_list, has_winning_line, legal_action_mask,
    portability_case_report, portability_report, validate_game_state, validate_move,
)
from .batch import (  # noqa: F401
    FEATURE_NAMES, SEED_DEFAULT, SEEDED_WEIGHTS, apply_moves, apply_symmetry, apply_symmetry_to_moves, bfs_level, book_level, canon_dedup, canonical_keys, canonical_payloads,
    decode_transform, encode_transform, eval_features, evaluate, guided_playouts, int_issue_peak, movegen_masks, opening_book_summary, orbit_sizes, payload_owner, perft, perft_probe,
    pipe_rates, playouts,
    random_roots, solve, solve_moves, symmetry_table, symmetry_table_rows, win_status,
)
from .engines import GpuRolloutEvaluator, batched_beam_engine, batched_mcts_engine, gpu_default_evaluate, gpu_simulate  # noqa: F401
from ._native import QuantikB200Error, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
