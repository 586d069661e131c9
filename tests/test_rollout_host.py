"""Host-side (CPU) checks of the rollout adapter's trajectory arithmetic."""


def test_trajectory_order_and_returns_small_case():
    """discounted_returns / trajectory_order on a hand-checked case."""
    import torch
    from safelife_b200.rollout import discounted_returns, trajectory_order
    dev = "cpu"
    done = torch.tensor([[0, 1], [0, 0], [1, 0]], dtype=torch.bool, device=dev)
    r = torch.tensor([[1., 2.], [3., 4.], [5., 6.]], dtype=torch.float64, device=dev)
    v = torch.tensor([[.5, .25], [1., 2.], [4., 8.]], dtype=torch.float64, device=dev)
    fv = torch.tensor([0., 10.], dtype=torch.float64, device=dev)          # env 0 ended, env 1 goes on
    g, lm = 0.5, 0.25
    ret, adv = discounted_returns(r, v, done, fv, g, lm)
    # env 0: one trajectory of 3 steps ending in done
    assert torch.allclose(ret[:, 0], torch.tensor([1 + .5 * (3 + .5 * 5), 3 + .5 * 5, 5.], dtype=torch.float64, device=dev))
    a2 = 5 - 4.
    a1 = 3 + .5 * 4 - 1 + lm * a2
    a0 = 1 + .5 * 1 - .5 + lm * a1
    assert torch.allclose(adv[:, 0], torch.tensor([a0, a1, a2], dtype=torch.float64, device=dev))
    # env 1: a one-step trajectory (done at t=0), then a running one bootstrapped with 10
    assert torch.allclose(ret[:, 1], torch.tensor([2., 4 + .5 * (6 + .5 * 10), 6 + .5 * 10], dtype=torch.float64, device=dev))
    b2 = 6 + .5 * 10 - 8
    b1 = 4 + .5 * 8 - 2 + lm * b2
    assert torch.allclose(adv[:, 1], torch.tensor([2 - .25, b1, b2], dtype=torch.float64, device=dev))
    # flat order: trajectories by (start step, env): (0,e0): t0,t1,t2; (0,e1): t0; (1,e1): t1,t2
    assert trajectory_order(done).tolist() == [0, 2, 4, 1, 3, 5]
