"""
The rollout adapter (safelife_b200/rollout.py) against a batch recorded from the
unmodified reference's PPO.gen_training_batch (training/ppo.py:75-143) over
reference envs under the training wrapper stack, with a fixed integer policy
(tests/golden/make_golden_ppo.py).  Observations, actions and the flat ordering
are exact; returns / advantages / values are float32 results of float32-float64
arithmetic on both sides: rtol 1e-5.
"""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _policy(d, device):
    import torch
    wp = torch.from_numpy(d["wp"]).to(device)
    wv = torch.from_numpy(d["wv"]).to(device)

    def policy(obs):
        counts = obs.to(torch.int64).sum(dim=(1, 2))
        # (integer matmul is not implemented on CUDA: broadcast and sum, still exact)
        action = torch.argmax((counts.unsqueeze(2) * wp.unsqueeze(0)).sum(dim=1), dim=1)
        policies = torch.nn.functional.one_hot(action, 9).to(torch.float32)
        return (counts * wv.unsqueeze(0)).sum(dim=1).to(torch.float32) / 64, policies
    return policy


@pytest.mark.parametrize("auto_reset", [True, False])
@pytest.mark.parametrize("kernel", [0, 1])
def test_gen_training_batch_matches_reference_ppo(auto_reset, kernel):
    import torch
    from safelife_b200 import wrappers
    from safelife_b200.env import BatchedSafeLifeEnv
    from safelife_b200.rollout import BatchedRollout
    d = gu.load_npz("ppo_batch.npz")
    levels = gu.levels_of(d)
    n, T = int(d["n_envs"]), int(d["steps"])
    base = BatchedSafeLifeEnv(levels, n, auto_reset=auto_reset, kernel=kernel,
                              time_limit=int(d["time_limit"]),
                              view_shape=tuple(int(v) for v in d["view_shape"]),
                              output_channels=tuple(int(c) for c in d["channels"]))
    env = wrappers.training_stack(base, exit_difficulty=float(d["exit_difficulty"]),
                                  se_penalty=float(d["se_penalty"]))
    runner = BatchedRollout(env)
    batch = runner.gen_training_batch(_policy(d, base.device), T, gamma=float(d["gamma"]),
                                      lmda=float(d["lmda"]),
                                      sample=lambda p: torch.argmax(p, dim=1))
    assert runner.num_steps == int(d["num_steps"])
    assert batch.obs.dtype == torch.uint8 and batch.actions.dtype == torch.int64
    assert np.array_equal(batch.actions.cpu().numpy(), d["batch_actions"])
    assert np.array_equal(batch.obs.cpu().numpy(), d["batch_obs"])
    assert np.array_equal(batch.action_prob.cpu().numpy(), d["batch_action_prob"])
    for name in ("values", "returns", "advantages"):
        got, want = getattr(batch, name).cpu().numpy(), d["batch_" + name]
        assert got.dtype == np.float32
        assert np.allclose(got, want, rtol=1e-5, atol=1e-5), (name, np.abs(got - want).max())
    # episodes did end inside the batch (trajectory stitching was exercised)
    assert int(runner.num_resets.sum()) > n
    # a second batch continues where the first one stopped
    again = runner.gen_training_batch(_policy(d, base.device), 5, sample=lambda p: torch.argmax(p, dim=1))
    assert again.obs.shape[0] == 5 * n
