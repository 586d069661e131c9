"""
Board-steps/s of the simulation part of side_effect_score (SURVEY 8 a10/a11,
config 5's episode-end measure) on one GPU (the reference C extension needs about 18 us per
board-step at 26x26 on one host core).

    python tools/bench_side_effects.py [pool] [boards] [num_samples]

One "board-step" = one advance_board step of one board (with the occupancy
histogram update where the reference does one).  A finished episode costs
num_steps + 2 * num_samples of them (side_effects.py:101-111).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    pos = [a for a in sys.argv[1:] if not a.startswith("--")]
    pool = pos[0] if len(pos) > 0 else "append-spawn-25"
    n = int(pos[1]) if len(pos) > 1 else 1024
    ns = int(pos[2]) if len(pos) > 2 else 1000
    import torch
    from safelife_b200 import side_effects as se
    from safelife_b200.env import seed_rng_states
    from safelife_b200.levels import load_levels, spawn_threshold
    levels = load_levels(os.path.join(ROOT, "tests", "golden", "pools", pool + ".npz"))
    rs = np.random.default_rng(0)
    init = np.stack([np.asarray(levels[i % len(levels)]["board"], dtype=np.uint16) for i in range(n)])
    final = init.copy()
    final[rs.random(init.shape) < 0.02] = 9 | (2 << 9)          # the agent's leftovers
    steps = rs.integers(200, 1000, n).astype(np.int32)
    thr = np.array([spawn_threshold(levels[i % len(levels)]["spawn_prob"]) for i in range(n)], dtype=np.int64)
    dev = torch.device("cuda", 0)
    t_init, t_final = (torch.from_numpy(a.view(np.int16)).to(dev) for a in (init, final))
    t_steps, t_thr = torch.from_numpy(steps).to(dev), torch.from_numpy(thr).to(dev)
    rng = torch.from_numpy(seed_rng_states(9, n).view(np.int64)).to(dev)
    se.occupancy_counts(t_init[:8], t_final[:8], t_steps[:8], t_thr[:8], rng[:8].clone(), 10)   # warm-up
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    counts = se.occupancy_counts(t_init, t_final, t_steps, t_thr, rng, ns)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    board_steps = int(steps.sum()) + 2 * ns * n
    out = {"pool": pool, "boards": n, "shape": list(init.shape[1:]), "num_samples": ns,
           "gpu_ms": ms, "gpu_board_steps_per_s": board_steps / (ms * 1e-3),
           "occupied": int((counts.sum(dim=(1, 2, 3, 4)) > 0).sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
