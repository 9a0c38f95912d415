"""Host-side logic that needs no GPU: module API / state_dict contract, install(), no-CPU-fallback guarantees."""
import copy
import sys
import types

import numpy as np
import pytest
import torch

import cswin_unet_b200 as cw
from cswin_unet_b200 import synth
from oracle import cswin_oracle as O
from tests import golden_util as G


def test_state_dict_keys_shapes_match_reference():
    z = G.load("model_t224")
    m = cw.cswin_tiny_224()
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in z["keys"]]
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in z["key_shapes"]]
    # strict load of a reference-shaped state dict
    shapes = O.state_dict_shapes()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, 1).items()}, strict=True)


def test_constructor_signatures_and_attributes():
    a = cw.LePEAttention(32, resolution=56, idx=0, split_size=1, num_heads=1)
    assert (a.H_sp, a.W_sp, a.resolution, a.num_heads) == (56, 1, 56, 1) and abs(a.scale - 32 ** -0.5) < 1e-12
    a = cw.LePEAttention(32, resolution=56, idx=1, split_size=2, num_heads=1, qk_scale=0.5)
    assert (a.H_sp, a.W_sp, a.scale) == (2, 56, 0.5)
    with pytest.raises(ValueError):
        cw.LePEAttention(32, resolution=56, idx=2, split_size=1)
    b = cw.CSWinBlock(dim=64, reso=56, num_heads=2, split_size=1, qkv_bias=True, drop_path=0.1)
    assert b.branch_num == 2 and b.patches_resolution == 56 and b.mlp_ratio == 4.0 and len(b.attns) == 2
    assert cw.CSWinBlock(dim=512, reso=7, num_heads=16, split_size=7).branch_num == 1      # reso == split => last stage
    assert sorted(k for k, _ in b.named_parameters()) == sorted(
        ["qkv.weight", "qkv.bias", "norm1.weight", "norm1.bias", "proj.weight", "proj.bias",
         "attns.0.get_v.weight", "attns.0.get_v.bias", "attns.1.get_v.weight", "attns.1.get_v.bias",
         "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias", "norm2.weight", "norm2.bias"])


def test_deepcopy_and_pickle_keep_parameters_and_drop_caches():
    import pickle
    b = cw.CSWinBlock(dim=64, reso=56, num_heads=2, split_size=1, qkv_bias=True)
    c = copy.deepcopy(b)
    assert c._derived is not b._derived
    assert all(torch.equal(p, q) and p.data_ptr() != q.data_ptr() for p, q in zip(b.parameters(), c.parameters()))
    d = pickle.loads(pickle.dumps(b))
    assert list(d.state_dict().keys()) == list(b.state_dict().keys())


def test_no_cpu_fallback():
    m = cw.cswin_tiny_224().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 224, 224))
    a = cw.LePEAttention(32, resolution=8, idx=0, split_size=2, num_heads=1).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        a(torch.zeros(3, 1, 64, 32))


def test_wrong_token_count_raises_like_the_reference():
    a = cw.LePEAttention(32, resolution=8, idx=0, split_size=2, num_heads=1)
    with pytest.raises(AssertionError, match="flatten img_tokens has wrong size"):
        a(torch.zeros(3, 1, 63, 32))


def test_install_rebinds_reference_module_globals():
    fake = types.ModuleType("fake_ref_networks_cswin_unet")
    for n in cw.install.__globals__["HOT_PATH_CLASSES"]:
        setattr(fake, n, object())
    sys.modules[fake.__name__] = fake
    saved = cw.install(fake.__name__)
    assert fake.LePEAttention is cw.LePEAttention and fake.CSWinBlock is cw.CSWinBlock and fake.CARAFE4 is cw.CARAFE4
    cw.uninstall(saved, fake.__name__)
    assert fake.LePEAttention is saved["LePEAttention"]


def test_droppath_matches_timm_semantics():
    torch.manual_seed(3)
    dp = cw.DropPath(0.25).train()
    x = torch.ones(8, 5, 4)
    s = dp.sample_scale(x)
    torch.manual_seed(3)
    m = x.new_empty(8, 1, 1).bernoulli_(0.75).div_(0.75)
    assert torch.equal(s, m.reshape(-1))
    assert dp.eval().sample_scale(x) is None


def test_synth_is_deterministic_and_name_keyed():
    a = synth.synth_tensor("stage1.0.qkv.weight", (192, 64), 7)
    b = synth.synth_tensor("stage1.0.qkv.weight", (192, 64), 7)
    c = synth.synth_tensor("stage1.0.proj.weight", (192, 64), 7)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(float(a.std()) - 1 / 8) < 0.01


def test_drop_path_plan_batches_the_draws_of_a_step():
    """modules.DropPathPlan: the first step records the call sequence (modules draw for themselves), later steps hand out rows of
    one (calls, B) Bernoulli draw with timm's scaling (mask / keep_prob); a changed call sequence falls back to per-call draws."""
    import torch
    from cswin_unet_b200 import modules
    plan = modules.DropPathPlan()
    dps = [modules.DropPath(p).train() for p in (0.1, 0.5, 0.0, 0.3)]
    x = torch.zeros(4096, 3, 2)
    try:
        for step in range(3):
            plan.begin(x.shape[0], "cpu")
            modules.DROP_PATH_PLAN = plan
            outs = [d.sample_scale(x) for d in dps]
            plan.end()
            modules.DROP_PATH_PLAN = None
            assert outs[2] is None                                            # p == 0: inactive, not part of the plan
            for d, o in zip(dps, outs):
                if o is None:
                    continue
                keep = 1.0 - d.drop_prob
                assert o.shape == (4096,) and o.dtype == torch.float32
                vals = set(round(v, 5) for v in o.unique().tolist())
                assert vals <= {0.0, round(1.0 / keep, 5)}
                assert abs((o > 0).float().mean().item() - keep) < 0.04       # Bernoulli(keep)
            if step > 0:
                assert plan.rows is not None and plan.rows.shape == (3, 4096) and plan.i == 3
        # a different call sequence: the plan declines, the module draws for itself
        plan.begin(x.shape[0], "cpu")
        modules.DROP_PATH_PLAN = plan
        o = modules.DropPath(0.25).train().sample_scale(x)
        assert o is not None and abs((o > 0).float().mean().item() - 0.75) < 0.04
    finally:
        modules.DROP_PATH_PLAN = None


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference path, rank 0 only) prints exactly one JSON line with the keys
    of the bench contract; library chatter on stdout is diverted to stderr."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "slices/s" and d["value"] > 0 and d["vs_baseline"] is None
    from baseline import ref_loader
    want_kind = "reference" if ref_loader.available() else "port"       # the unmodified reference when baseline/_ref travelled
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # other ranks of a torchrun launch exit 0 without work and without output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    o2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                        text=True, timeout=600, cwd=root, env=env)
    assert o2.returncode == 0 and o2.stdout.strip() == ""


def test_bench_native_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: the native arm must fail loudly where there is no CUDA device (this container)."""
    import os, subprocess, sys
    import torch
    if torch.cuda.is_available():
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode != 0 and "CUDA" in (out.stderr + out.stdout)
