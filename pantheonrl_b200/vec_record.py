"""Recorded trajectories from the vectorised rollout buffers ("on device").

With n_envs = 1 the reference records by wrapping the env (`recorder_wrap`, pantheonrl/common/wrappers.py:84-232).
The device engine already holds everything a recorder would have seen — the ego's `[T][N]` buffer and the
partner's ragged `[Tcap][N]` buffer — so the same `TurnBasedTransitions` /
`SimultaneousTransitions` objects (and `.npy` files) can be cut out of them after the fact, one env at a time.
Host-side numpy on copies of the buffers (`tensor.cpu().numpy()`): nothing here is on the hot path.

Order of a turn-based env's moves: rows are merged by (episode, move number inside the episode); the episode of
a row is the running count of its buffer's `episode_starts`, the move number comes from `move_index(obs)`
(Liar's Dice: the number of bids already on the table).  Valid for buffers that start at an episode boundary
(the first rollout, or any rollout of a game whose episodes all end inside a tick, e.g. RPS).
"""
import numpy as np

from .common.trajsaver import SimultaneousTransitions, TurnBasedTransitions
from .common.wrappers import ALT_DONE, ALT_NOT_DONE, DONE, EGO_DONE, EGO_NOT_DONE, NOT_DONE


def liar_move_index(obs):
    """Bids on the table = history pairs that are not padding (liar.py:54-56 pads with face 6)."""
    o = np.asarray(obs)
    return int((o[..., 6:30:2] != 6).sum(axis=-1))


def turn_based_transitions(ego, alt, env, last_done, obs_len=30, act_len=2, move_index=liar_move_index):
    """`TurnBasedTransitions` of env `env`: what a `TurnBasedRecorder` around that env would hold.
    ego / alt: dicts of host arrays (obs [T][N][32], actions [T][N][4], episode_starts [T][N], alt['count'] [N]);
    last_done: whether the last ego step of the buffer ended its episode (carry.ego_last_done[env])."""
    T = ego["obs"].shape[0]
    K = int(alt["count"][env])
    rows = []
    ep = 0
    for t in range(T):
        ep += int(ego["episode_starts"][t, env] != 0)
        o = ego["obs"][t, env, :obs_len]
        rows.append((ep, move_index(o), 0, t, o, ego["actions"][t, env, :act_len]))
    ep = 0
    for j in range(K):
        ep += int(alt["episode_starts"][j, env] != 0)
        o = alt["obs"][j, env, :obs_len]
        rows.append((ep, move_index(o), 1, j, o, alt["actions"][j, env, :act_len]))
    rows.sort(key=lambda r: (r[0], r[1]))
    # an episode is complete if it ended before the ego's last row, or with it (last_done); a partner opening
    # move after the final auto-reset belongs to an episode that has only just begun
    ego_last_ep = max((r[0] for r in rows if r[2] == 0), default=0)
    flags = []
    for i, r in enumerate(rows):
        ends = i + 1 == len(rows) or rows[i + 1][0] != r[0]
        complete = r[0] < ego_last_ep or (r[0] == ego_last_ep and bool(last_done))
        done = ends and complete
        flags.append((EGO_DONE if done else EGO_NOT_DONE) if r[2] == 0 else (ALT_DONE if done else ALT_NOT_DONE))
    obs = np.array([r[4] for r in rows]).astype(np.int64).reshape(len(rows), obs_len)
    acts = np.array([r[5] for r in rows]).astype(np.int64).reshape(len(rows), act_len)
    return TurnBasedTransitions(obs, acts, np.array(flags))


def simultaneous_transitions(ego, alt, env, last_done, obs_len, act_len=1):
    """`SimultaneousTransitions` of env `env` (one partner row per tick: RPS, Overcooked)."""
    T = ego["obs"].shape[0]
    assert int(alt["count"][env]) == T, "a simultaneous game stores one partner row per tick"
    starts = ego["episode_starts"][:, env]
    flags = np.array([(DONE if (starts[t + 1] if t + 1 < T else last_done) else NOT_DONE) for t in range(T)])
    cut = lambda a, n: np.asarray(a[:T, env, :n]).astype(np.int64 if a.dtype == np.uint8 else a.dtype)  # noqa: E731
    egoacts, altacts = cut(ego["actions"], act_len), cut(alt["actions"], act_len)
    if act_len == 1:  # Discrete actions are stored flat by the reference's recorder
        egoacts, altacts = egoacts.reshape(-1), altacts.reshape(-1)
    return SimultaneousTransitions(cut(ego["obs"], obs_len), egoacts, cut(alt["obs"], obs_len), altacts, flags)
