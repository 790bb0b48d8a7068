"""Which random numbers the N = 1 facade consumes.

"philox" (default): counter-based Philox streams on the device: dice, coins, action
    samples and shuffles are functions of (seed, stream, index, tick) — reproducible, shardable.
"reference": the reference's own three global generators, drawn on the host at the points and in
    the order the reference draws them, and handed to the kernels:
      * np.random — the who-starts coin (multiagentenv.py:325), 12 dice per reset (liar.py:22-26,
        97-100), partner resampling (multiagentenv.py:115) and SB3's per-epoch
        np.random.permutation (RolloutBuffer.get);
      * torch's default CPU generator — weight init (policy.py) and every action sample:
        Categorical.sample() = torch.multinomial(probs, 1) = argmax(probs / Exp(1) draws), one
        exponential per logit, heads in order; the host draws them, pth_policy_forward runs the race;
      * set_random_seed(seed) semantics of every PPO constructor: random, np.random and torch are all
        re-seeded (trainer.py:111-112, 198-199 -> SB3 BaseAlgorithm).
    With it a run of `trainer.py LiarsDice-v0 PPO PPO --seed 10` shaped code sees the dice, coins
    and — logits permitting — the actions the reference would see on a CPU.
Only the host-driven flow (n_envs = 1) has this mode; the device engine always uses Philox.
"""
_MODE = "philox"


def set_rng_mode(mode):
    global _MODE
    if mode not in ("philox", "reference"):
        raise ValueError("rng mode is 'philox' or 'reference'")
    _MODE = mode


def get_rng_mode():
    return _MODE
