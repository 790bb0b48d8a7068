"""Reference-RNG mode on the GPU (pantheonrl_b200/rng_mode.py): with the host drawing from the
reference's generators at the reference's points, the device-backed LiarEnv replays the reference's own
game trace, pth_policy_forward's exponential race equals torch.multinomial on the same generator state,
and PPO.train consumes np.random.permutation like SB3's RolloutBuffer.get."""
import os

import numpy as np
import pytest
import torch

from pantheonrl_b200 import ops
from pantheonrl_b200.common.agents import Agent, OnPolicyAgent
from pantheonrl_b200.envs import LiarEnv, RPSEnv
from pantheonrl_b200.ppo import PPO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "compat_liar.npz")


class Scripted(Agent):
    def __init__(self, acts):
        self.acts, self.k = acts, 0

    def get_action(self, obs, record=True):
        a = self.acts[self.k]
        self.k += 1
        return a

    def update(self, reward, done):
        pass


def test_liar_env_replays_the_reference_trace_under_np_seed_10(ctx):
    g = np.load(GOLD)
    np.random.seed(10)
    env = LiarEnv(rng="reference")
    env.add_partner_agent(Scripted(g["alt_acts"].astype(np.int64)))
    obs = env.reset()
    T = g["obs"].shape[0]
    assert T == 2048
    for t in range(T):
        assert np.array_equal(np.asarray(obs), g["obs"][t]), t
        obs, r, d, _ = env.step(g["ego_acts"][t].astype(np.int64))
        assert r == g["rew"][t] and int(d) == g["done"][t], t
        if d:
            obs = env.reset()
    assert np.array_equal(np.asarray(obs), g["final_obs"]) and env.episodes == len(g["resets"])
    assert np.random.randint(1 << 30) == g["np_state_after"][0]


@pytest.mark.parametrize("make_env", [RPSEnv, LiarEnv])
def test_race_sampling_equals_torch_multinomial(ctx, make_env):
    env = make_env()
    model = PPO("MlpPolicy", env, seed=10, rng="reference")
    model.policy.params.mul_(40.0)  # far from uniform: the argmax is not decided by the noise alone
    rs = np.random.RandomState(0)
    sp = env.observation_space
    nvec = sp.nvec if hasattr(sp, "nvec") else np.array([sp.n])
    torch.manual_seed(77)
    seen = set()
    for k in range(300):
        obs = np.array([rs.randint(n) for n in nvec])
        state = torch.get_rng_state()
        actions, values, logp = model.policy.forward(obs)
        after = torch.get_rng_state()
        # the same draw through torch's own sampling code on OUR logits
        model.policy._stage_obs(obs)
        out = ops.policy_forward(model.space, model.policy.params, model.policy._obs_dev,
                                 action_in=torch.zeros(1, 4, dtype=torch.uint8, device="cuda"), want=("logits",))
        logits = out["logits"].cpu()
        torch.set_rng_state(state)
        want, o = [], 0
        for n in model.space.heads:
            dist = torch.distributions.Categorical(logits=logits[:, o:o + n])
            want.append(int(dist.sample()))
            o += n
        assert torch.equal(torch.get_rng_state(), after), "the facade drew a different number of random numbers"
        assert np.asarray(actions).reshape(-1).tolist() == want, (k, actions, want)
        seen.add(tuple(want))
    assert len(seen) > 2


def test_reference_mode_training_consumes_numpy_permutations(ctx):
    env = RPSEnv()
    partner = OnPolicyAgent(PPO("MlpPolicy", env, n_steps=64, batch_size=16, n_epochs=3, seed=10, rng="reference"))
    env.add_partner_agent(partner)
    ego = PPO("MlpPolicy", env, n_steps=64, batch_size=16, n_epochs=3, seed=10, rng="reference")
    assert torch.equal(ego.policy.params, partner.model.policy.params)  # both constructors re-seed (trainer.py:111, 198)
    ego.learn(total_timesteps=64)
    # RPS resets draw nothing; the only np.random consumers were the ego's three epoch permutations
    np.random.seed(10)
    want = np.stack([np.random.permutation(64) for _ in range(3)])
    assert np.array_equal(ego._perm.cpu().numpy(), want)
    r = ego.rollout_buffer.h["rewards"]
    assert 0.5 < float((r != 0).mean()) < 0.8  # the two learners do not share their samples


def test_adap_reference_mode_draws_from_torch_like_the_reference(ctx):
    """ADAP in the reference-RNG mode: the first context is drawn right after the policy is built
    (adap_learn.py:212-217), and train() draws, per minibatch, th.randperm(B) and then num_context_samples
    contexts (adap/util.py:106, 113-114) from torch's generator — nothing else touches it."""
    from pantheonrl_b200.adap import ADAP, AdapPolicy, _torch_sampler
    from pantheonrl_b200 import policy as pol
    env = RPSEnv()
    model = ADAP(AdapPolicy, env, seed=10, rng="reference", n_steps=128, batch_size=64, n_epochs=2,
                 num_context_samples=4, num_state_samples=16)
    torch.manual_seed(10)
    pol.init_flat(model.space, None, 3)                      # the constructor's weight draws
    want_ctx = _torch_sampler("l2", 3, 1)                    # then the first context
    assert torch.equal(model.policy.get_context(), want_ctx)
    # a filled buffer (any numbers), then train(): compare the generator state with a replay of the reference's draws
    buf = model.rollout_buffer
    rs = np.random.RandomState(0)
    for _ in range(128):
        row = np.concatenate(([0.0], rs.randn(3)))
        buf.add(row, np.array([rs.randint(3)]), float(rs.randn()), False, torch.tensor([0.1]), torch.tensor([-1.1]))
    buf.compute_returns_and_advantage(torch.tensor([0.0]), False)
    torch.manual_seed(123)
    np.random.seed(5)
    model.train()
    after = torch.get_rng_state()
    np_after = np.random.randint(1 << 30)
    torch.manual_seed(123)
    for _ in range(2 * 2):                                   # n_epochs x minibatches
        torch.randperm(64)
        for _ in range(4):
            _torch_sampler("l2", 3, 1)
    assert torch.equal(torch.get_rng_state(), after)
    cl = model.last_context_loss.cpu().numpy()
    assert cl.shape == (4,) and np.all((cl > 0) & (cl <= 1 + 1e-6))
    np.random.seed(5)  # SB3's RolloutBuffer.get: one np.random.permutation per epoch
    for _ in range(2):
        np.random.permutation(128)
    assert np.random.randint(1 << 30) == np_after
