"""golden for evaluation.features / evaluate and the eval-guided rollout policy (run with the reference)"""
import sys, os, json, time
import numpy as np
REF = os.environ.get('QUANTIK_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, 'src')); sys.path.insert(0, HERE)
from quantik_core.evaluation import features, evaluate, EvalConfig
from quantik_core.game_utils import WinStatus, check_game_winner, has_winning_line
from quantik_core.move import apply_move, generate_legal_moves, generate_legal_moves_list
import gen_golden as GG
M32 = GG.M32
def guided_playout(root, seed, rollout, root_index, eps, cfg, max_plies=-1):
    bb = tuple(int(x) for x in root); ply = 0
    while True:
        if max_plies >= 0 and ply >= max_plies: return 0, ply
        w = check_game_winner(bb)
        if w != WinStatus.NO_WIN: return (1 if w == WinStatus.PLAYER_0_WINS else -1), ply
        player, by = generate_legal_moves(bb)
        moves = [m for ms in by.values() for m in ms]
        if not moves: return (-1 if player == 0 else 1), ply
        r = GG.philox4x32_10((rollout & M32, (rollout >> 32) & M32, 0x80000000 | (ply // 2), root_index), (seed & M32, (seed >> 32) & M32))
        r_eps, r_pick = r[2 * (ply % 2)], r[2 * (ply % 2) + 1]
        # MCTSEngine._select_rollout_move (mcts.py:345-383) with the random module replaced by the stream
        if eps > 0 and r_eps * (1.0 / 4294967296.0) < eps:
            move = moves[(r_pick * len(moves)) >> 32]
        else:
            best_move, best_score = moves[0], float("-inf")
            move = None
            for m in moves:
                child = apply_move(bb, m)
                if has_winning_line(child) or not generate_legal_moves_list(child):
                    move = m; break
                sc = evaluate(child, player, cfg)
                if sc > best_score: best_score, best_move = sc, m
            if move is None: move = best_move
        bb = apply_move(bb, move); ply += 1
d = np.load(os.path.join(HERE, '..', 'tests', 'golden') + '/states_reference.npz')
nv = int(d['n_valid'])
st = d['states'][:nv][d['win'][:nv] == 0][:1200]
F0 = np.array([features(tuple(int(x) for x in s), 0) for s in st]); F1 = np.array([features(tuple(int(x) for x in s), 1) for s in st])
fitted = EvalConfig.load()          # tuning/weights.json of the source checkout
seeded = EvalConfig()
print('fitted', fitted.weights, 'seeded', seeded.weights)
E0f = np.array([evaluate(tuple(int(x) for x in s), 0, fitted) for s in st])
out = {'states': st, 'features_p0': F0, 'features_p1': F1, 'fitted_weights': fitted.weights, 'seeded_weights': seeded.weights, 'evaluate_p0_fitted': E0f}
seed = GG.SEED
t = time.time()
empty = (0,) * 8
for name, cfg, eps in (('seeded_eps02', seeded, 0.2), ('fitted_eps02', fitted, 0.2), ('fitted_eps0', fitted, 0.0), ('seeded_eps1', seeded, 1.0)):
    res = [guided_playout(empty, seed, j, 3, eps, cfg) for j in range(160)]
    out[name + '_outcome'] = np.array([a for a, _ in res], np.int8); out[name + '_plies'] = np.array([b for _, b in res], np.uint8)
    print(name, sum(a > 0 for a, _ in res), np.mean([b for _, b in res]), time.time() - t, flush=True)
p = np.load(os.path.join(HERE, '..', 'tests', 'golden') + '/playouts_philox.npz')
roots = p['roots'][:12]
moc = np.zeros((len(roots), 24), np.int8); mpl = np.zeros((len(roots), 24), np.uint8)
for r, bb in enumerate(roots):
    for j in range(24):
        moc[r, j], mpl[r, j] = guided_playout(bb, seed, 500 + j, 40 + r, 0.2, fitted)
out['roots'] = roots; out['roots_outcome'] = moc; out['roots_plies'] = mpl
res = [guided_playout(empty, seed, j, 3, 0.2, fitted, 9) for j in range(80)]
out['cut9_outcome'] = np.array([a for a, _ in res], np.int8); out['cut9_plies'] = np.array([b for _, b in res], np.uint8)
np.savez_compressed(os.path.join(HERE, '..', 'tests', 'golden') + '/guided_rollouts.npz', **out)
print('done', time.time() - t)
