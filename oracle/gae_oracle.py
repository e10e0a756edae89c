"""TEST INFRASTRUCTURE ONLY -- numpy restatement of Stable-Baselines3 2.6.0 ``RolloutBuffer.compute_returns_and_advantage``
(the learner the reference drives from ballbot_rl/training/train.py:126-141; SB3 is a third-party dependency pinned in
the reference's requirements.txt:12 and absent from this container, so the published algorithm is restated here).

SB3 walks the buffer backwards with ``next_non_terminal = 1 - episode_starts[step + 1]`` (``1 - dones`` for the last
step) -- episode_starts[t + 1] is exactly the done flag of the transition taken at step t.
"""
import numpy as np


def gae(rewards, values, dones, last_values, gamma=0.99, gae_lambda=0.95):
    """rewards/values/dones [T,N]; last_values [N] -> advantages, returns (float32 like SB3's buffers)."""
    T, N = rewards.shape
    rewards = rewards.astype(np.float32); values = values.astype(np.float32)
    adv = np.zeros((T, N), np.float32)
    last = np.zeros(N, np.float32)
    for step in reversed(range(T)):
        next_values = last_values.astype(np.float32) if step == T - 1 else values[step + 1]
        nonterminal = 1.0 - dones[step].astype(np.float32)
        delta = rewards[step] + np.float32(gamma) * next_values * nonterminal - values[step]
        last = delta + np.float32(gamma) * np.float32(gae_lambda) * nonterminal * last
        adv[step] = last
    return adv, adv + values
