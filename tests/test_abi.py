"""CPU tests of the C-ABI boundary: the shared library loads and exports every
symbol include/pantheon_b200.h declares; no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pantheon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pth_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_symbols()
    for must in ("pth_gae_f32", "pth_env_rps_step", "pth_env_liar_step", "pth_policy_forward",
                 "pth_rollout_run", "pth_ppo_update", "pth_pack_transitions", "pth_version"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build_cuda()
    from pantheonrl_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert sorted(_lib.SIGNATURES) == names, set(names) ^ set(_lib.SIGNATURES)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.pth_version() == 100


def test_struct_sizes_match_the_header():
    from pantheonrl_b200 import _lib
    assert ctypes.sizeof(_lib.Space) == 4 * (2 + 96 + 1 + 4)
    assert ctypes.sizeof(_lib.Buffer) == 8 * 8
    assert ctypes.sizeof(_lib.EnvCarry) == 9 * 8
    assert ctypes.sizeof(_lib.OvercookedLayout) == 5576 and _lib.PTH_OC_STATE_BYTES == 40
    assert ctypes.sizeof(_lib.RolloutArgs) % 8 == 0 and _lib.RolloutArgs.d_layout.offset % 8 == 0


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pantheonrl_b200 import _lib
    with pytest.raises(_lib.PthError):
        _lib.Context(0)
    # host-only helpers still work without a device
    lib = _lib.load()
    sp = _lib.Space.onehot([7] * 6 + [7, 12] * 12, [7, 12])
    assert lib.pth_space_feature_dim(ctypes.byref(sp)) == 270
    assert lib.pth_space_logit_dim(ctypes.byref(sp)) == 19
    assert lib.pth_policy_param_count(ctypes.byref(sp)) == 44308
    assert lib.pth_policy_param_count(ctypes.byref(_lib.Space.onehot([1], [3]))) == 8836
    assert lib.pth_policy_param_count(ctypes.byref(_lib.Space.box(62, [6]))) == 16839


def test_ctypes_mirrors_have_the_size_gcc_gives_the_header_structs(tmp_path):
    """The header compiles as plain C, and every ctypes.Structure in _lib.py is as large as the struct it
    mirrors (a field added on one side only would shift every later argument silently)."""
    import subprocess
    from pantheonrl_b200 import _lib
    pairs = [("pth_space", "Space"), ("pth_update_args", "UpdateArgs"), ("pth_forward_args", "ForwardArgs"),
             ("pth_rollout_args", "RolloutArgs"), ("pth_buffer", "Buffer"), ("pth_env_carry", "EnvCarry"),
             ("pth_overcooked_layout", "OvercookedLayout")]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "pantheon_b200.h"\nint main(void) {\n' +
                   "".join(f'  printf("%zu\\n", sizeof({c}));\n' for c, _ in pairs) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    for (c, py), n in zip(pairs, sizes):
        assert ctypes.sizeof(getattr(_lib, py)) == n, (c, py, n)
