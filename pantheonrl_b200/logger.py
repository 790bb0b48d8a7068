"""SB3-shaped scalar logger.

The reference logs through stable-baselines3's ``Logger``:
``configure_logger(verbose, tensorboard_log, tb_log_name)`` in
``OnPolicyAgent.__init__`` (pantheonrl/common/agents.py:102-103), then
``model.logger.record(key, value, exclude=...)`` / ``model.logger.dump(step=...)``
(agents.py:132-153; SB3 ``PPO.train`` restated at
pantheonrl/algos/adap/adap_learn.py:354-371, ``learn`` at :487-500).  The website
reads the TensorBoard scalars back by tag (website/data_processing.py:211-221),
so tags and steps are the contract: ``rollout/ep_rew_mean``, ``rollout/ep_len_mean``,
``time/{iterations,fps,time_elapsed,total_timesteps}``, ``train/{entropy_loss,
policy_gradient_loss, value_loss, approx_kl, clip_fraction, loss, explained_variance,
n_updates, clip_range}``.

Host-side plumbing only: the numbers are produced on the device (update kernel
statistics, rollout episode counters) and arrive here as Python floats.
"""
import os
import sys
from collections import OrderedDict


class HumanOutputFormat:
    """SB3's stdout table: keys grouped by their ``section/`` prefix."""

    def __init__(self, stream=None, max_length=36):
        self.stream, self.max_length = stream or sys.stdout, max_length

    def write(self, key_values, key_excluded, step=0):
        rows, section = [], None
        for key in sorted(key_values):
            if "stdout" in key_excluded.get(key, ()) or "log" in key_excluded.get(key, ()):
                continue
            value = key_values[key]
            text = f"{value:<8.3g}" if isinstance(value, float) else str(value)
            if "/" in key:
                sec, name = key.split("/", 1)
                if sec != section:
                    rows.append((sec + "/", ""))
                    section = sec
                rows.append(("   " + name, text))
            else:
                section = None
                rows.append((key, text))
        if not rows:
            return
        rows = [(self._trunc(k), self._trunc(v)) for k, v in rows]
        kw, vw = max(len(k) for k, _ in rows), max(len(v) for _, v in rows)
        dashes = "-" * (kw + vw + 7)
        lines = [dashes] + [f"| {k:<{kw}} | {v:<{vw}} |" for k, v in rows] + [dashes]
        self.stream.write("\n".join(lines) + "\n")
        self.stream.flush()

    def _trunc(self, s):
        return s if len(s) <= self.max_length else s[: self.max_length - 3] + "..."

    def close(self):
        pass


class CSVOutputFormat:
    """progress.csv like SB3's (new keys extend the header; the file is rewritten then)."""

    def __init__(self, filename):
        self.filename, self.keys, self.rows = filename, [], []

    def write(self, key_values, key_excluded, step=0):
        kv = {k: v for k, v in key_values.items() if "csv" not in key_excluded.get(k, ())}
        new = [k for k in kv if k not in self.keys]
        self.keys += new
        self.rows.append(kv)
        with open(self.filename, "w") as f:
            f.write(",".join(self.keys) + "\n")
            for r in self.rows:
                f.write(",".join("" if r.get(k) is None else str(r[k]) for k in self.keys) + "\n")

    def close(self):
        pass


class TensorBoardOutputFormat:
    """Scalars by tag and step, the way SB3 writes them (strings are skipped)."""

    def __init__(self, folder):
        from torch.utils.tensorboard import SummaryWriter
        self.writer = SummaryWriter(log_dir=folder)

    def write(self, key_values, key_excluded, step=0):
        for key, value in key_values.items():
            if "tensorboard" in key_excluded.get(key, ()) or isinstance(value, str) or value is None:
                continue
            self.writer.add_scalar(key, float(value), step)
        self.writer.flush()

    def close(self):
        self.writer.close()


class Logger:
    """``record`` / ``record_mean`` / ``dump`` with per-key format exclusions."""

    def __init__(self, folder=None, output_formats=()):
        self.dir, self.output_formats = folder, list(output_formats)
        self.name_to_value, self.name_to_count, self.name_to_excluded = OrderedDict(), {}, {}

    @staticmethod
    def _excl(exclude):
        if exclude is None:
            return ()
        return (exclude,) if isinstance(exclude, str) else tuple(exclude)

    def record(self, key, value, exclude=None):
        self.name_to_value[key] = value
        self.name_to_excluded[key] = self._excl(exclude)

    def record_mean(self, key, value, exclude=None):
        if value is None:
            return
        n = self.name_to_count.get(key, 0)
        old = self.name_to_value.get(key, 0.0)
        self.name_to_value[key] = old * n / (n + 1) + value / (n + 1)
        self.name_to_count[key] = n + 1
        self.name_to_excluded[key] = self._excl(exclude)

    def dump(self, step=0):
        for fmt in self.output_formats:
            fmt.write(self.name_to_value, self.name_to_excluded, step)
        self.name_to_value.clear()
        self.name_to_count.clear()
        self.name_to_excluded.clear()

    def get_dir(self):
        return self.dir

    def close(self):
        for fmt in self.output_formats:
            fmt.close()


def get_latest_run_id(log_path, log_name):
    """Highest N among ``{log_path}/{log_name}_N`` (SB3 utils.get_latest_run_id)."""
    best = 0
    if log_path and os.path.isdir(log_path):
        for entry in os.listdir(log_path):
            head, _, tail = entry.rpartition("_")
            if head == log_name and tail.isdigit():
                best = max(best, int(tail))
    return best


def configure_logger(verbose=0, tensorboard_log=None, tb_log_name="", reset_num_timesteps=True):
    """SB3 ``utils.configure_logger``: a new ``{tensorboard_log}/{tb_log_name}_{N+1}`` run
    directory with TensorBoard (+ stdout when verbose), stdout only when just verbose,
    nothing otherwise."""
    folder, formats = None, []
    if tensorboard_log is not None:
        run = get_latest_run_id(tensorboard_log, tb_log_name) + (1 if reset_num_timesteps else 0)
        folder = os.path.join(tensorboard_log, f"{tb_log_name}_{max(run, 1)}")
        os.makedirs(folder, exist_ok=True)
        formats.append(TensorBoardOutputFormat(folder))
    if verbose >= 1:
        formats.insert(0, HumanOutputFormat())
    return Logger(folder, formats)


def safe_mean(values):
    """SB3 ``safe_mean``: nan for an empty list instead of a warning."""
    values = list(values)
    return float("nan") if not values else float(sum(values)) / len(values)


def explained_variance(values, returns):
    """1 - Var[returns - values] / Var[returns] (SB3 ``explained_variance``), on whatever
    device the tensors live; nan when the returns are constant."""
    import torch
    y_pred, y_true = values.reshape(-1).float(), returns.reshape(-1).float()
    var_y = torch.var(y_true, unbiased=False)
    if float(var_y) == 0.0:
        return float("nan")
    return float(1.0 - torch.var(y_true - y_pred, unbiased=False) / var_y)
