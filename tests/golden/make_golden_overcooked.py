"""Generate the Overcooked golden fixtures by RUNNING THE REFERENCE's own code.

Run in the authoring container only:  python tests/golden/make_golden_overcooked.py
Imports, verbatim and unmodified (through tests/golden/ref_shim.py stubs for gym / SB3):
  /root/reference/overcookedgym/overcooked.py                        (OvercookedMultiEnv)
  /root/reference/pantheonrl/common/multiagentenv.py                 (SimultaneousEnv routing)
  /root/reference/overcookedgym/human_aware_rl/overcooked_ai/overcooked_ai_py/{mdp,planning,agents}
The fixtures are committed; tests never read /root/reference.

Fixtures
  oc_ref_featurization.npz  THE REFERENCE'S OWN GOLDEN VECTOR
        overcooked_ai_py/data/testing/state_featurization.pickle (checked by the reference's
        overcooked_test.py:169-173): 5 games x 400 states x 2 players x 62 features on layout
        `simple` under GreedyHumanModel pairs after np.random.seed(0).  The pickle only holds
        features, so the generating run is repeated here (the live run reproduces the pickle
        bit for bit - asserted below) and the joint actions + rewards are stored beside it.
  oc_full_traj.npz          THE REFERENCE'S OWN GOLDEN TRAJECTORY
        overcooked_ai/common_tests/trajectory_tests/test_full_traj.json (checked by
        overcooked_test.py:135-142): 30 transitions on layout `mdp_test` (tomatoes, cook_time 5,
        order list), re-encoded as arrays (canonical state bytes, see encode_state).
  oc_layouts.npz            grids + parameters of every layout trainer.py accepts (LAYOUT_LIST)
  oc_routing_<layout>_e<idx>.npz
        OvercookedMultiEnv driven through MultiAgentEnv.step / reset with a scripted ego and a
        recording scripted partner (epsilon-greedy GreedyHumanModel): everything the ego and the
        partner see (obs, rewards, dones, update calls) across episode boundaries.
  oc_random_<layout>.npz    uniformly random joint actions through OvercookedMultiEnv.multi_step
        (covers counter drops / pick-ups and collisions the greedy agents never produce)
"""
import json
import os
import pickle
import sys

import numpy as np

np.Inf = np.inf  # overcooked_env.py:132 uses the NumPy-1 alias

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
OA = os.path.join(ref_shim.REF, "overcookedgym/human_aware_rl/overcooked_ai")
sys.path.insert(0, OA)

from overcooked_ai_py.agents.agent import AgentPair, GreedyHumanModel  # noqa: E402
from overcooked_ai_py.mdp.actions import Action, Direction  # noqa: E402
from overcooked_ai_py.mdp.overcooked_env import OvercookedEnv  # noqa: E402
from overcooked_ai_py.mdp.overcooked_mdp import OvercookedGridworld, OvercookedState, PlayerState, ObjectState  # noqa: E402
from overcooked_ai_py.planning import planners  # noqa: E402
from overcooked_ai_py.planning.planners import NO_COUNTERS_PARAMS, MediumLevelPlanner  # noqa: E402

# The reference would pickle the planner next to its sources (planners.py:927-934); /root/reference
# is read-only, so compute it in memory instead.  Nothing else is patched.
_MLP_CACHE = {}


def _mlp_in_memory(mdp, mlp_params, custom_filename=None, force_compute=False):
    key = (mdp.layout_name, json.dumps(mlp_params, sort_keys=True, default=str))
    if key not in _MLP_CACHE:
        _MLP_CACHE[key] = MediumLevelPlanner(mdp, mlp_params)
    return _MLP_CACHE[key]


MediumLevelPlanner.from_pickle_or_compute = staticmethod(_mlp_in_memory)

from overcookedgym.overcooked import OvercookedMultiEnv  # noqa: E402
from overcookedgym.overcooked_utils import LAYOUT_LIST  # noqa: E402
from pantheonrl.common.agents import Agent  # noqa: E402

OBJ = {"onion": 1, "soup": 2, "dish": 3, "tomato": 4}
SOUP_TYPE = {"onion": 0, "tomato": 1}
ORDER = {"onion": 0, "tomato": 1, "any": 2}


def aidx(a):
    return Action.ACTION_TO_INDEX[a if a == "interact" else tuple(a)]


def encode_state(state, W, H):
    """Canonical bytes: per player (x, y, orientation index, held, soup type, soup n, soup t),
    per cell y*W+x (object, soup type, soup n, soup t), order list length (255 = None) + entries."""
    out = []
    for p in state.players:
        o = p.held_object
        row = [p.position[0], p.position[1], Direction.DIRECTION_TO_INDEX[tuple(p.orientation)], 0, 0, 0, 0]
        if o is not None:
            row[3] = OBJ[o.name]
            if o.name == "soup":
                row[4:7] = [SOUP_TYPE[o.state[0]], o.state[1], o.state[2]]
        out += row
    cells = np.zeros((H * W, 4), np.uint8)
    for pos, o in state.objects.items():
        c = pos[1] * W + pos[0]
        cells[c, 0] = OBJ[o.name]
        if o.name == "soup":
            cells[c, 1:4] = [SOUP_TYPE[o.state[0]], o.state[1], o.state[2]]
    ol = state.order_list
    order = [255] + [0] * 8 if ol is None else [len(ol)] + [ORDER[x] for x in ol] + [0] * (8 - len(ol))
    return np.concatenate([np.array(out, np.uint8), cells.reshape(-1), np.array(order, np.uint8)])


def gen_ref_featurization():
    mdp = OvercookedGridworld.from_layout_name("simple")
    mlp = MediumLevelPlanner(mdp, NO_COUNTERS_PARAMS)
    env = OvercookedEnv(mdp, horizon=400)
    pair = AgentPair(GreedyHumanModel(mlp), GreedyHumanModel(mlp))
    np.random.seed(0)
    # == OvercookedEnv.get_rollouts(pair, num_games=5) (overcooked_env.py:162-205) with the
    # run_agents loop (overcooked_env.py:131-160) written out: its final np.array(trajectory)
    # of ragged tuples no longer works on NumPy 2; every call into the reference is unchanged.
    acts, sparse, shaped, states = [], [], [], []
    for _ in range(5):
        pair.set_mdp(env.mdp)
        done = False
        ea, es, eh, st = [], [], [], []
        while not done:
            s_t = env.state
            a_t = pair.joint_action(s_t)
            _, r, done, info = env.step(a_t)
            st.append(s_t)
            ea.append([aidx(a_t[0]), aidx(a_t[1])])
            es.append(r)
            eh.append(info["shaped_r"])
        acts.append(ea); sparse.append(es); shaped.append(eh); states.append(st)
        env.reset()
        pair.reset()
    live = np.array([[mdp.featurize_state(s, mlp) for s in ep] for ep in states])
    with open(os.path.join(OA, "overcooked_ai_py/data/testing/state_featurization.pickle"), "rb") as f:
        expected = np.array(pickle.load(f))
    assert expected.shape == (5, 400, 2, 62) and np.array_equal(expected, live), \
        "the live reference run no longer reproduces the reference's golden pickle"
    assert np.array_equal(expected, expected.astype(np.int8))
    np.savez_compressed(os.path.join(HERE, "oc_ref_featurization.npz"),
                        feats=expected.astype(np.int8), actions=np.array(acts, np.uint8),
                        sparse=np.array(sparse, np.int16), shaped=np.array(shaped, np.int16))


def gen_full_traj():
    path = os.path.join(OA, "common_tests/trajectory_tests/test_full_traj.json")
    d = json.load(open(path))
    mp = d["mdp_params"][0]
    mdp = OvercookedGridworld.from_layout_name(**mp)
    H, W = len(mdp.terrain_mtx), len(mdp.terrain_mtx[0])
    states = []
    for s in d["ep_observations"][0]:
        players = []
        for p in s["players"]:
            ho = p["held_object"]
            obj = None if ho is None else ObjectState(ho["name"], tuple(ho["position"]),
                                                      None if ho["state"] is None else tuple(ho["state"]))
            players.append(PlayerState(tuple(p["position"]), tuple(p["orientation"]), obj))
        objs = {tuple(o["position"]): ObjectState(o["name"], tuple(o["position"]),
                                                  None if o["state"] is None else tuple(o["state"]))
                for o in s["objects"]}
        states.append(OvercookedState(players, objs, order_list=s["order_list"]))
    acts = [[aidx(a[0]), aidx(a[1])] for a in d["ep_actions"][0]]
    # the reference's own check (AgentEvaluator.check_trajectories): replaying reproduces the file
    for i in range(len(states) - 1):
        ja = tuple(Action.INDEX_TO_ACTION[k] for k in acts[i])
        nxt, r, _ = mdp.get_state_transition(states[i], ja)
        assert nxt == states[i + 1] and r == d["ep_rewards"][0][i]
    ja = tuple(Action.INDEX_TO_ACTION[k] for k in acts[-1])
    last, r_last, _ = mdp.get_state_transition(states[-1], ja)
    assert r_last == d["ep_rewards"][0][-1]
    enc = np.stack([encode_state(s, W, H) for s in states + [last]])
    np.savez_compressed(os.path.join(HERE, "oc_full_traj.npz"),
                        grid=np.array(["".join(r) for r in mdp.terrain_mtx]),
                        start=np.array(mdp.start_player_positions, np.int32),
                        cook_time=mp["cook_time"], num_items=mp["num_items_for_soup"],
                        delivery_reward=mdp.delivery_reward,
                        order_list=np.array([ORDER[x] for x in mp["start_order_list"]], np.uint8),
                        actions=np.array(acts, np.uint8), rewards=np.array(d["ep_rewards"][0], np.int16),
                        states=enc)


def gen_layouts():
    out = {}
    for name in LAYOUT_LIST:
        mdp = OvercookedGridworld.from_layout_name(name)
        out[name] = dict(grid=["".join(r) for r in mdp.terrain_mtx],
                         start=[list(p) for p in mdp.start_player_positions],
                         cook_time=mdp.soup_cooking_time, num_items=mdp.num_items_for_soup,
                         delivery_reward=mdp.delivery_reward)
    np.savez_compressed(os.path.join(HERE, "oc_layouts.npz"), layouts=json.dumps(out))
    return out


class EpsGreedy:
    """epsilon-greedy GreedyHumanModel for player `idx` of the base env (test driver only)."""

    def __init__(self, env, idx, eps, seed):
        mlp = _mlp_in_memory(env.mdp, NO_COUNTERS_PARAMS)
        self.env, self.idx, self.eps = env, idx, eps
        self.rng = np.random.RandomState(seed)
        self.agent = GreedyHumanModel(mlp)
        self.agent.set_agent_index(idx)
        self.agent.set_mdp(env.mdp)

    def __call__(self):
        if self.rng.rand() < self.eps:
            return int(self.rng.randint(6))
        try:
            return aidx(self.agent.action(self.env.base_env.state))
        except Exception:  # noqa: BLE001  (greedy model has no plan in some layouts)
            return int(self.rng.randint(6))


class ScriptedPartner(Agent):
    def __init__(self, draw, log):
        self.draw, self.log = draw, log

    def get_action(self, obs, record=True):
        a = self.draw()
        self.log.append(("act", np.array(obs.obs).copy(), a))
        return a

    def update(self, reward, done):
        self.log.append(("upd", float(reward), bool(done)))


def gen_routing(layout, ego_idx, T, eps, seed, horizon_patch=None):
    np.random.seed(seed)
    env = OvercookedMultiEnv(layout, ego_agent_idx=ego_idx)
    if horizon_patch:
        env.base_env.horizon = horizon_patch  # shorter episodes -> more reset boundaries per fixture
    log = []
    env.add_partner_agent(ScriptedPartner(EpsGreedy(env, 1 - ego_idx, eps, seed + 1), log))
    ego = EpsGreedy(env, ego_idx, eps, seed + 2)
    ego_obs, ego_act, ego_rew, ego_done, ev_at = [], [], [], [], [0]
    obs = env.reset()
    for _ in range(T):
        a = ego()
        o2, r, d, info = env.step(a)
        ego_obs.append(np.array(obs).copy()); ego_act.append(a); ego_rew.append(float(r)); ego_done.append(int(d))
        obs = env.reset() if d else o2  # DummyVecEnv auto-reset
        ev_at.append(len(log))
    kind = np.array([0 if e[0] == "act" else 1 for e in log], np.uint8)
    ev_obs = np.zeros((len(log), 62), np.int8)
    ev_act = np.zeros(len(log), np.uint8)
    ev_rew = np.zeros(len(log), np.float32)
    ev_done = np.zeros(len(log), np.uint8)
    for i, e in enumerate(log):
        if e[0] == "act":
            assert np.array_equal(e[1], e[1].astype(np.int8))
            ev_obs[i], ev_act[i] = e[1], e[2]
        else:
            ev_rew[i], ev_done[i] = e[1], e[2]
    eo = np.array(ego_obs)
    assert np.array_equal(eo, eo.astype(np.int8))
    np.savez_compressed(os.path.join(HERE, f"oc_routing_{layout}_e{ego_idx}.npz"),
                        layout=layout, ego_idx=ego_idx, horizon=env.base_env.horizon,
                        ego_obs=eo.astype(np.int8), ego_act=np.array(ego_act, np.uint8),
                        ego_rew=np.array(ego_rew, np.float32), ego_done=np.array(ego_done, np.uint8),
                        ev_at_step=np.array(ev_at, np.int32), ev_kind=kind, ev_obs=ev_obs, ev_act=ev_act,
                        ev_rew=ev_rew, ev_done=ev_done, final_obs=np.array(obs).astype(np.int8))
    return float(np.sum(ego_rew))


def gen_random(layout, T, seed, p_interact=0.3):
    rng = np.random.RandomState(seed)
    env = OvercookedMultiEnv(layout)
    W, H = len(env.mdp.terrain_mtx[0]), len(env.mdp.terrain_mtx)
    acts, obs0, obs1, rew, states, dones = [], [], [], [], [], []
    (o0, o1) = env.multi_reset()
    for _ in range(T):
        a = [5 if rng.rand() < p_interact else int(rng.randint(5)) for _ in range(2)]
        states.append(encode_state(env.base_env.state, W, H))
        obs0.append(o0); obs1.append(o1)
        (o0, o1), (r, _), d, _ = env.multi_step(a[0], a[1])
        acts.append(a); rew.append(r); dones.append(int(d))
        if d:  # horizon reached: the caller (MultiAgentEnv.reset) starts a new episode
            (o0, o1) = env.multi_reset()
    states.append(encode_state(env.base_env.state, W, H))
    obs0.append(o0); obs1.append(o1)
    np.savez_compressed(os.path.join(HERE, f"oc_random_{layout}.npz"), layout=layout,
                        actions=np.array(acts, np.uint8), obs0=np.array(obs0).astype(np.int8),
                        obs1=np.array(obs1).astype(np.int8), rewards=np.array(rew, np.int16),
                        dones=np.array(dones, np.uint8), states=np.stack(states))
    return int(np.sum(rew))


if __name__ == "__main__":
    gen_ref_featurization()
    gen_full_traj()
    gen_layouts()
    print("routing simple e0", gen_routing("simple", 0, 1300, 0.15, 5))
    print("routing simple e1", gen_routing("simple", 1, 600, 0.3, 6, horizon_patch=150))
    print("routing unident_s e0", gen_routing("unident_s", 0, 500, 0.2, 7))
    print("routing random0 e1", gen_routing("random0", 1, 500, 0.2, 8))
    for lay, seed in (("simple", 11), ("random1", 12), ("corridor", 13), ("scenario2_s", 14)):
        print("random", lay, gen_random(lay, 2000, seed))
