"""
Rollout adapter: the batched form of ``BaseAlgo.obs_for_envs`` /
``BaseAlgo.act_on_envs`` (training/base_algo.py:152-244), which in the
reference walk a Python list of envs, step them one by one and call
``env.reset()`` on the ones that finished.

    runner = BatchedRollout(env)                 # BatchedSafeLifeEnv or a wrapper stack over it
    obs, agent_ids = runner.obs_for_envs()
    next_obs, reward, done = runner.act_on_envs(actions)

Semantics kept from the reference:
  * the first ``obs_for_envs`` resets every env;
  * ``act_on_envs`` returns the observation of the step itself (the terminal
    observation for envs that finished), then resets finished envs onto their
    next level and remembers the fresh observation for the next
    ``obs_for_envs`` (base_algo.py:231-238);
  * agent ids are (env index, number of resets so far, agent index 0): the key
    PPO uses to stitch trajectories (training/ppo.py:93-120), here an int64
    [N, 3] tensor instead of a list of tuples.
Everything stays on the device; the env must be built with auto_reset=False.
"""
import torch


class BatchedRollout(object):
    def __init__(self, env):
        base = getattr(env, "unwrapped", env)
        if base.auto_reset:
            raise ValueError("BatchedRollout drives the resets itself: use auto_reset=False")
        self.env, self.base = env, base
        self.last_obs = None
        n, dev = base.num_envs, base.device
        self.num_resets = torch.zeros(n, dtype=torch.int64, device=dev)
        self._index = torch.arange(n, dtype=torch.int64, device=dev)

    def agent_ids(self):
        return torch.stack([self._index, self.num_resets, torch.zeros_like(self._index)], dim=1)

    def obs_for_envs(self):
        if self.last_obs is None:
            self.last_obs = self.env.reset().clone()
        return self.last_obs, self.agent_ids()

    def act_on_envs(self, actions):
        if self.last_obs is None:
            raise RuntimeError("call obs_for_envs() first")
        obs, reward, done, info = self.env.step(actions)
        next_obs = obs.clone()                    # terminal observation where done
        reward, done = reward.clone(), done.clone()
        self.last_info = info
        fresh = self.env.reset_done(done)
        self.num_resets += done.to(torch.int64)
        self.last_obs = fresh.clone()
        return next_obs, reward, done

    def take_one_step(self, policy):
        """obs -> policy(obs) -> actions -> step; mirrors BaseAlgo.take_one_step (:246-270)."""
        obs, agent_ids = self.obs_for_envs()
        actions = policy(obs)
        next_obs, rewards, done = self.act_on_envs(actions)
        return obs, actions, rewards, done, next_obs, agent_ids
