"""Generate golden fixtures by RUNNING THE REFERENCE's own Python classes.

Run in the authoring container only:  python tests/golden/make_golden.py
Imports /root/reference/pantheonrl/{envs/rpsgym/rps.py, envs/liargym/liar.py,
common/multiagentenv.py} verbatim (through tests/golden/ref_shim.py) and writes
small .npz fixtures next to this file.  The fixtures are committed; tests never
read /root/reference.

Fixtures
  rps_payoff.npz    RPSEnv.multi_step over the full 3x3 action table
  liar_env.npz      LiarEnv.multi_reset / ego_step / alt_step traces with random
                    (legal and illegal) raw actions
  routing_rps.npz   MultiAgentEnv.step/reset event traces at N=1 with a scripted
  routing_liar.npz  ego and recording scripted partner(s): every get_action /
                    update call the partner receives, every (obs, reward, done)
                    the ego receives, every reset's dice and who-starts coin
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from pantheonrl.common.agents import Agent  # noqa: E402
from pantheonrl.envs.liargym.liar import LiarEnv  # noqa: E402
from pantheonrl.envs.rpsgym.rps import RPSEnv  # noqa: E402


def gen_rps_payoff():
    env = RPSEnv()
    rows = []
    for a in range(3):
        for b in range(3):
            (o0, o1), (r0, r1), done, _ = env.multi_step(a, b)
            rows.append((a, b, r0, r1, int(done), int(o0[0]), int(o1[0])))
    np.savez(os.path.join(HERE, "rps_payoff.npz"), table=np.array(rows, np.int32))


def draw_liar_action(rng, last_count):
    """Raw action in MultiDiscrete([7, 12]); skewed so that games get long."""
    mode = rng.rand()
    if mode < 0.55 and last_count < 11:
        return np.array([rng.randint(6), min(11, last_count + 1 + (rng.rand() < 0.2))])
    return np.array([rng.randint(7), rng.randint(12)])


def gen_liar_env(n_episodes=400, seed=1234):
    rng = np.random.RandomState(seed)
    np.random.seed(seed + 1)  # the env's dice use the global generator (liar.py:25)
    env = LiarEnv()
    ep_start, hands, egofirst, first_obs = [], [], [], []
    is_ego, raw, obs_out, r_ego, r_alt, done_l = [], [], [], [], [], []
    for _ in range(n_episodes):
        ef = bool(rng.rand() < 0.5)
        o = env.multi_reset(ef)
        ep_start.append(len(raw))
        hands.append(list(env.egohand) + list(env.althand))
        egofirst.append(int(ef))
        first_obs.append(o)
        turn_ego = ef
        last_count = -1
        while True:
            a = draw_liar_action(rng, last_count)
            ob, (r0, r1), d, _ = env.ego_step(a) if turn_ego else env.alt_step(a)
            is_ego.append(int(turn_ego))
            raw.append(a)
            obs_out.append(ob)
            r_ego.append(r0)
            r_alt.append(r1)
            done_l.append(int(d))
            if d:
                break
            last_count = env.history[1]
            turn_ego = not turn_ego
    ep_start.append(len(raw))
    np.savez(os.path.join(HERE, "liar_env.npz"),
             ep_start=np.array(ep_start, np.int32), hands=np.array(hands, np.uint8),
             egofirst=np.array(egofirst, np.uint8), first_obs=np.array(first_obs, np.uint8),
             is_ego=np.array(is_ego, np.uint8), raw_action=np.array(raw, np.uint8),
             obs=np.array(obs_out, np.uint8), r_ego=np.array(r_ego, np.float32),
             r_alt=np.array(r_alt, np.float32), done=np.array(done_l, np.uint8))


class ScriptedPartner(Agent):
    """Returns pre-scripted actions and logs every call it receives."""

    def __init__(self, pid, rng, draw, log):
        self.pid, self.rng, self.draw, self.log = pid, rng, draw, log

    def get_action(self, obs, record=True):
        a = self.draw(self.rng, obs.obs)
        self.log.append(("act", self.pid, np.array(obs.obs).copy(), np.array(a).copy()))
        return a

    def update(self, reward, done):
        self.log.append(("upd", self.pid, float(reward), bool(done)))


def run_routing(env, n_partners, T, ego_draw, alt_draw, seed, get_reset_info, out):
    log = []
    for p in range(n_partners):
        env.add_partner_agent(ScriptedPartner(p, np.random.RandomState(seed + 10 + p), alt_draw, log))
    ego_rng = np.random.RandomState(seed + 5)
    np.random.seed(seed)
    ego_obs, ego_act, ego_rew, ego_done, ego_pid = [], [], [], [], []
    ev_at_step = [0]
    reset_info = []
    obs = env.reset()
    reset_info.append(get_reset_info(env))
    reset_pid = [env.partnerids[0]]
    for _ in range(T):
        a = ego_draw(ego_rng, obs)
        o2, r, d, info = env.step(a)
        ego_obs.append(np.array(obs).copy())
        ego_act.append(np.array(a).copy())
        ego_rew.append(float(r))
        ego_done.append(int(d))
        ego_pid.append(int(info["_partnerid"][0]))
        if d:
            obs = env.reset()  # DummyVecEnv auto-reset
            reset_info.append(get_reset_info(env))
            reset_pid.append(env.partnerids[0])
        else:
            obs = o2
        ev_at_step.append(len(log))
    # flatten the partner event log
    kind = np.array([0 if e[0] == "act" else 1 for e in log], np.uint8)
    pid = np.array([e[1] for e in log], np.int32)
    obs_dim = np.array(ego_obs[0]).size
    act_dim = np.array(ego_act[0]).size
    ev_obs = np.zeros((len(log), obs_dim), np.uint8)
    ev_act = np.zeros((len(log), act_dim), np.uint8)
    ev_rew = np.zeros(len(log), np.float32)
    ev_done = np.zeros(len(log), np.uint8)
    for i, e in enumerate(log):
        if e[0] == "act":
            ev_obs[i] = np.asarray(e[2]).reshape(-1)
            ev_act[i] = np.asarray(e[3]).reshape(-1)
        else:
            ev_rew[i] = e[2]
            ev_done[i] = e[3]
    np.savez(out, ego_obs=np.array(ego_obs, np.uint8).reshape(T, -1),
             ego_act=np.array(ego_act, np.uint8).reshape(T, -1),
             ego_rew=np.array(ego_rew, np.float32), ego_done=np.array(ego_done, np.uint8),
             ego_pid=np.array(ego_pid, np.int32), ev_at_step=np.array(ev_at_step, np.int32),
             ev_kind=kind, ev_pid=pid, ev_obs=ev_obs, ev_act=ev_act, ev_rew=ev_rew,
             ev_done=ev_done, reset_info=np.array(reset_info, np.uint8),
             reset_pid=np.array(reset_pid, np.int32), final_obs=np.array(obs, np.uint8).reshape(-1))


class LoggedLiar(LiarEnv):
    """LiarEnv that remembers the dice and coin of its latest multi_reset."""

    def multi_reset(self, egofirst):
        o = super().multi_reset(egofirst)
        self.last_reset = [int(egofirst)] + list(self.egohand) + list(self.althand)
        return o


def gen_routing():
    def liar_draw(rng, obs):
        obs = np.asarray(obs).reshape(-1)
        last_count = -1 if obs[6] == 6 and obs[7] == 0 and obs[8] == 6 else int(obs[7])
        if obs[6] == 6:  # empty history (DEFAULT pad)
            last_count = -1
        return draw_liar_action(rng, last_count)

    for n_partners, name in ((1, "routing_liar.npz"), (3, "routing_liar_rr3.npz")):
        env = LoggedLiar()
        run_routing(env, n_partners, 700, liar_draw, liar_draw, 77 + n_partners,
                    lambda e: list(e.last_reset), os.path.join(HERE, name))

    def rps_draw(rng, obs):
        return rng.randint(3)

    env = RPSEnv()
    run_routing(env, 1, 64, rps_draw, rps_draw, 99, lambda e: [0], os.path.join(HERE, "routing_rps.npz"))


if __name__ == "__main__":
    gen_rps_payoff()
    gen_liar_env()
    gen_routing()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
