"""
Rollout adapter: the batched form of the reference's PPO data collection --
``BaseAlgo.obs_for_envs`` / ``act_on_envs`` (training/base_algo.py:152-244) and
``PPO.take_one_step`` / ``PPO.gen_training_batch`` (training/ppo.py:61-143) --
which in the reference walk a Python list of envs, step them one by one, call
``env.reset()`` on the ones that finished and stitch per-agent trajectories in a
dict keyed by ``(id(env), num_resets, k)``.

    runner = BatchedRollout(env)                 # BatchedSafeLifeEnv or a wrapper stack over it
    batch = runner.gen_training_batch(policy, steps_per_env=20, gamma=0.97, lmda=0.95)
    # batch.obs, .actions, .action_prob, .returns, .advantages, .values   (flat, on the device)

`policy(obs) -> (values [n], policies [n, 9])` is the model call of
``PPO.take_one_step`` (ppo.py:64-66); `sample(policies) -> actions [n]` replaces
the per-env ``get_rng().choice`` (ppo.py:68; default: ``torch.multinomial``).

Semantics kept from the reference:
  * the first step resets every env (base_algo.py:167-172);
  * a step returns the observation of the step itself; a finished env is then
    reset and its fresh observation is what the policy sees next
    (base_algo.py:231-238).  With ``auto_reset=True`` (the fast path: one launch
    per step) the reset happens inside the step's launch and the returned
    observation of a finished env IS the fresh one; the reference never uses
    the terminal observation of a finished env (its value bootstrap is skipped
    when done, ppo.py:117-121);
  * agent ids are (env index, number of resets so far, agent 0);
  * trajectories end at `done`; the value of the observation after the last
    step bootstraps the trajectories that are still running (ppo.py:115-121);
  * returns / advantages by the reference's recursion (ppo.py:123-135:
    ``returns[i] += gamma * returns[i+1]``, ``advantages[i] += lmda *
    advantages[i+1]``), in float64, cast to float32 at the end (ppo.py:139-141);
  * the flat output is ordered like the reference's dict of trajectories: by the
    step at which a trajectory started, then by env, then by time.
Everything stays on the device: observations are written by the step kernel
straight into the rollout buffer.
"""
import collections

import torch

TrainingBatch = collections.namedtuple(
    "TrainingBatch", "obs actions action_prob returns advantages values")
StepData = collections.namedtuple(
    "StepData", "obs actions rewards done next_obs agent_ids policies values")


class BatchedRollout(object):
    def __init__(self, env):
        self.env = env
        self.base = getattr(env, "unwrapped", env)
        n, dev = self.base.num_envs, self.base.device
        self.last_obs = None
        self.num_resets = torch.zeros(n, dtype=torch.int64, device=dev)
        self._index = torch.arange(n, dtype=torch.int64, device=dev)
        self.num_steps = 0

    # ---- BaseAlgo.obs_for_envs / act_on_envs
    def agent_ids(self):
        return torch.stack([self._index, self.num_resets, torch.zeros_like(self._index)], dim=1)

    def obs_for_envs(self):
        if self.last_obs is None:
            self.last_obs = self.env.reset()
        return self.last_obs, self.agent_ids()

    def act_on_envs(self, actions, obs_out=None):
        """
        One step of every env.  Returns (next_obs, reward, done): with auto_reset the
        observation of a finished env is the first of its next episode (one launch);
        without, it is the terminal one and the reset is a second call, exactly
        base_algo.py:231-238.
        """
        if self.last_obs is None:
            raise RuntimeError("call obs_for_envs() first")
        obs, reward, done, info = self.env.step(actions, obs_out=obs_out)
        self.last_info = info
        if self.base.auto_reset:
            self.last_obs = obs
        else:
            self.last_obs = self.env.reset_done(done)
        self.num_resets = self.num_resets + done.to(torch.int64)
        return obs, reward, done

    def take_one_step(self, policy, sample=None, obs_out=None):
        """PPO.take_one_step (ppo.py:61-73) for the whole batch."""
        obs, agent_ids = self.obs_for_envs()
        values, policies = policy(obs)
        actions = (sample or _multinomial)(policies)
        next_obs, rewards, done = self.act_on_envs(actions, obs_out=obs_out)
        return StepData(obs, actions, rewards, done, next_obs, agent_ids, policies, values)

    # ---- PPO.gen_training_batch
    def gen_training_batch(self, policy, steps_per_env, gamma=0.97, lmda=0.95, sample=None,
                           reference_order=True):
        """
        PPO.gen_training_batch (ppo.py:75-143): `steps_per_env` steps of every env,
        trajectories cut at episode ends, discounted returns and advantages.
        Returns flat device tensors (obs in the env's dtype; the rest float32 /
        int64 as the reference's ``t(label)``).
        """
        assert steps_per_env > 0
        T, N, dev = int(steps_per_env), self.base.num_envs, self.base.device
        first_obs, _ = self.obs_for_envs()
        # observations of steps 0..T: slot t is what the policy saw at step t; the step
        # kernel writes slot t + 1 directly
        obs_buf = torch.empty((T + 1,) + tuple(first_obs.shape), dtype=first_obs.dtype, device=dev)
        obs_buf[0].copy_(first_obs)
        self.last_obs = obs_buf[0]
        actions = torch.empty((T, N), dtype=torch.int64, device=dev)
        action_prob = torch.empty((T, N), dtype=torch.float32, device=dev)
        rewards = torch.empty((T, N), dtype=torch.float64, device=dev)
        values = torch.empty((T, N), dtype=torch.float64, device=dev)
        done = torch.empty((T, N), dtype=torch.bool, device=dev)
        with torch.no_grad():
            for t in range(T):
                out = obs_buf[t + 1] if self.base.auto_reset else None
                step = self.take_one_step(policy, sample, obs_out=out)
                a = step.actions.to(torch.int64)
                actions[t] = a
                action_prob[t] = step.policies.gather(1, a.unsqueeze(1)).squeeze(1)
                rewards[t] = step.rewards
                values[t] = step.values.reshape(N)
                done[t] = step.done
                if not self.base.auto_reset:
                    obs_buf[t + 1].copy_(self.last_obs)
                    self.last_obs = obs_buf[t + 1]
                    final_obs = step.next_obs
            # value of the observation after the last step, for the trajectories that go on
            # (next_obs of the last step where not done: equal to the next policy input)
            final_value = policy(obs_buf[T] if self.base.auto_reset else final_obs)[0].reshape(N)
            final_value = final_value.to(torch.float64) * (~done[T - 1])
        returns, advantages = discounted_returns(rewards, values, done, final_value, gamma, lmda)
        self.num_steps += T * N
        self.last_obs = obs_buf[T].clone()      # the buffer is handed to the caller
        order = trajectory_order(done) if reference_order else None

        def flat(x, dtype=None):
            x = x.reshape((T * N,) + tuple(x.shape[2:]))
            if order is not None:
                x = x[order]
            return x if dtype is None else x.to(dtype)
        return TrainingBatch(flat(obs_buf[:T]), flat(actions), flat(action_prob),
                             flat(returns, torch.float32), flat(advantages, torch.float32),
                             flat(values, torch.float32))


def _multinomial(policies):
    return torch.multinomial(policies, 1).squeeze(1)


def discounted_returns(rewards, values, done, final_value, gamma, lmda):
    """
    ppo.py:123-135 on [T, N] tensors (float64): per trajectory
        val1 = values shifted by one, final_value at the end
        advantages = rewards + gamma * val1 - val0;  returns[-1] += gamma * final_value
        returns[i] += gamma * returns[i+1];  advantages[i] += lmda * advantages[i+1]
    where a trajectory ends at `done` (final value 0) or at the end of the batch
    (final value = `final_value`, already zeroed where the last step was done).
    """
    T = rewards.shape[0]
    returns = torch.empty_like(rewards)
    advantages = torch.empty_like(rewards)
    next_value, next_return = final_value, gamma * 0 + final_value
    next_adv = torch.zeros_like(final_value)
    for t in range(T - 1, -1, -1):
        last = done[t] if t < T - 1 else torch.ones_like(done[t])    # last step of its trajectory
        v1 = torch.where(done[t], torch.zeros_like(next_value), next_value)
        delta = rewards[t] + gamma * v1 - values[t]
        if t == T - 1:
            returns[t] = rewards[t] + gamma * v1
            advantages[t] = delta
        else:
            returns[t] = rewards[t] + gamma * torch.where(last, torch.zeros_like(next_return), next_return)
            advantages[t] = delta + lmda * torch.where(last, torch.zeros_like(next_adv), next_adv)
        next_value, next_return, next_adv = values[t], returns[t], advantages[t]
    return returns, advantages


def trajectory_order(done):
    """
    Permutation of the flat [T * N] index (t * N + e) that lists the steps the way the
    reference's ``defaultdict`` of trajectories does (ppo.py:88-111, 137-141):
    trajectories in order of first appearance -- by starting step, then env -- and
    each trajectory's steps in time order.
    """
    T, N = done.shape
    dev = done.device
    t_idx = torch.arange(T, device=dev).unsqueeze(1).expand(T, N)
    # start step of the trajectory of (t, e): one past the last done before t
    started = torch.zeros((T, N), dtype=torch.int64, device=dev)
    if T > 1:
        after_done = torch.where(done[:-1], t_idx[1:], torch.zeros_like(t_idx[1:]))
        started[1:] = torch.cummax(after_done, dim=0).values
    e_idx = torch.arange(N, device=dev).unsqueeze(0).expand(T, N)
    key = (started * N + e_idx) * T + t_idx
    return torch.argsort(key.reshape(-1))
