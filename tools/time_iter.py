"""Quick per-phase timing of VecTrainer iterations (dev tool)."""
import sys, os, json, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pantheonrl_b200.engine import VecTrainer, PPOConfig


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(env, N, T, partner, iters=4, nmb=32, epochs=10, **kw):
    cfg = PPOConfig(n_steps=T, n_minibatches=nmb, n_epochs=epochs)
    tr = VecTrainer(env, N, cfg, seed=10, partner=partner, **kw)
    out = []
    for it in range(iters):
        e = [ev() for _ in range(4)]
        e[0].record(); tr.collect(); e[1].record(); tr.compute_gae(); e[2].record(); m = tr.train(); e[3].record()
        torch.cuda.synchronize()
        out.append(dict(rollout_ms=e[0].elapsed_time(e[1]), gae_ms=e[1].elapsed_time(e[2]),
                        train_ms=e[2].elapsed_time(e[3]), alt_M=m))
    st = tr.train_stats()
    es = tr.episode_stats()
    steps = N * T + (out[-1]["alt_M"] if partner == "ppo" else N * T)
    tot = out[-1]["rollout_ms"] + out[-1]["gae_ms"] + out[-1]["train_ms"]
    print(env, N, T, partner, json.dumps(out[-1]), f"agent-steps/s={steps / tot * 1e3:.3e}")
    print("   ", {k: round(v, 5) if isinstance(v, float) else v for k, v in st.items()}, es)


if __name__ == "__main__":
    run("liar", 4096, 128, "ppo")
    run("overcooked", 1024, 400, "ppo", layout="simple")
    run("overcooked", 4096, 400, "ppo", layout="simple", iters=3)
    run("rps", 65536, 128, "selfplay")
    run("rps", 4096, 128, "ppo")
    run("liar", 1, 2048, "ppo", iters=2, nmb=0)
