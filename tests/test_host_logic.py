"""Host-side logic that needs no GPU: config, parameter table, generators, the C-ABI surface, the
reference-facing module, frame sharding over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import PKG, ROOT
from fvp import capi, config as fcfg, dist as fdist, netspec, synth


def test_presets_and_yaml_merge(tmp_path):
    cfg = fcfg.preset("campus")
    assert cfg.DATASET.CAMERA_NUM == 3 and cfg.CAPTURE_SPEC.MAX_PEOPLE == 5 and cfg.DATASET.NUM_JOINTS == 17
    y = tmp_path / "c.yaml"
    y.write_text("DATASET:\n  CAMERA_NUM: 4\nCAPTURE_SPEC:\n  MIN_SCORE: 0.2\n")
    c2 = fcfg.load_yaml(str(y))
    assert c2.DATASET.CAMERA_NUM == 4 and c2.CAPTURE_SPEC.MIN_SCORE == 0.2 and c2.DATASET.NUM_JOINTS == 15
    y.write_text("NO_SUCH_KEY: 1\n")
    with pytest.raises(ValueError):          # like lib/core/config.py:187-188
        fcfg.load_yaml(str(y))


def test_param_table_counts():
    rows = netspec.param_table(15)
    assert len(rows) == 485                                            # SURVEY.md section 5 "checkpoint"
    by_net = lambda p: sum(1 for k, _, _ in rows if k.startswith(p))
    assert (by_net("pose_net.center_net"), by_net("pose_net.c2c_net"), by_net("joint_net.conv_net"),
            by_net("joint_net.weight_net")) == (162, 156, 156, 11)
    n_params = sum(int(np.prod(s)) for k, s, d in rows if d == "float32" and "running" not in k)
    assert n_params == 2636788
    assert netspec.macs_per_image(netspec.p2p_net(15), (64, 64)) == 620560384          # exact count of the layer table
    assert netspec.macs_per_image(netspec.c2c_net(15), (20,)) == 2471360              # 2.5 MMAC per column (App. B)


def test_conv_mac_counts_match_survey():
    # App. B: P2PNet 620.6 MMAC per 64x64 image, CenterNet 1085 MMAC at 80x80 (J=15)
    assert abs(netspec.macs_per_image(netspec.p2p_net(15), (64, 64)) / 1e6 - 620.6) < 0.5
    assert abs(netspec.macs_per_image(netspec.center_net(15), (80, 80)) / 1e6 - 1085) < 2


def test_weight_generator_is_deterministic(golden):
    g = golden("campus_b1")        # Golden() already asserts the SHA-256 of the regenerated weights
    again = synth.make_weights(17, seed=int(g["weight_seed"]))
    assert all(np.array_equal(again[k], g.weights[k]) for k in again)


def test_resize_transform_closed_form(golden):
    for name in ("panoptic_b2", "campus_b1", "shelf_crowd"):
        g = golden(name)
        A = synth.resize_transform(g.cfg.DATASET.ORI_IMAGE_SIZE, g.cfg.DATASET.IMAGE_SIZE)
        np.testing.assert_allclose(A, g.resize, atol=1e-4, rtol=1e-6)


def test_heatmap_lattice_roundtrip(golden):
    g = golden("panoptic_none_valid")
    q = synth.quantise_u16(g.heatmaps)
    assert np.array_equal(synth.dequantise_u16(q), g.heatmaps)
    assert g.heatmaps.max() <= 1.0 and g.heatmaps.min() >= 0.0


# ---- C ABI ---------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(built_library):
    """Every function include/fvp_b200.h declares is exported and bound (no compute, no GPU)."""
    header = open(os.path.join(ROOT, "include", "fvp_b200.h")).read()
    declared = set(re.findall(r"\b(fvp_[a-z0-9_]+)\s*\(", header))
    declared -= {"fvp_ctx", "fvp_config"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.fvp_abi_version() == capi.ABI_VERSION
    assert ctypes.sizeof(capi.FvpConfig) == 4 * (4 + 4 + 3 + 3 + 3 + 3 + 3 + 1 + 1 + 1 + 2 + 2)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(capi.FvpLibraryError):
        capi.load(str(tmp_path / "libfvp_b200.so"))


def test_library_is_sm100a_with_lineinfo(built_library):
    out = subprocess.run(["cuobjdump", "-lelf", built_library], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_bench_conv_tensor_rooflines():
    """bench.py's tensor-pipe figures: useful FLOPs = 2 x reference MACs, executed = 3 x (hi/lo split), against a given peak."""
    import bench
    st = {"p2p_net": 0.352, "center_net": 0.178}
    r = bench.conv_tensor_rooflines(st, 1, 10, 15, (80, 80), 1635.6)
    assert abs(r["p2p_net"]["gmac"] - 18.618) < 0.01 and abs(r["center_net"]["gmac"] - 1.085) < 0.002
    assert abs(r["p2p_net"]["useful_tflops"] - 2 * 18.618 / 0.352) < 0.1           # GMAC / ms = TMAC/s
    assert abs(r["p2p_net"]["frac_executed"] - 3 * r["p2p_net"]["useful_tflops"] / 1635.6) < 1e-12
    r32 = bench.conv_tensor_rooflines({"p2p_net": 6.38, "center_net": 0.65}, 32, 10, 15, (80, 80), 1635.6)
    assert abs(r32["p2p_net"]["useful_tflops"] - 2 * 32 * 18.618 / 6.38) < 0.5


def _conv_plan(lib, n, H, W, cin, cin2, cout, k, engine=2, sms=148):
    import ctypes as C
    out = (C.c_int * 10)()
    assert lib.fvp_debug_conv_plan(n, H, W, cin, cin2, cout, k, engine, sms, out) == 0
    keys = ("n_tile", "n_tiles", "resident", "a_stages", "b_stages", "smem", "occ2", "grid", "items", "tmem")
    return dict(zip(keys, list(out)))


def test_conv_launch_plans(built_library):
    """Host-side planner of the tcgen05 convolution (N-tile width, weight residency, CTAs per SM), queried without a GPU:
    resource invariants over every trunk layer shape x batch size, and the decisions DESIGN.md 4.1 describes."""
    from fvp import capi
    lib = capi.load()
    layers = [(16, 0, 16, 7), (16, 0, 32, 3), (32, 16, 32, 3), (32, 0, 32, 3), (32, 0, 64, 3), (64, 32, 64, 3), (64, 0, 64, 3),
              (64, 0, 128, 3), (128, 64, 128, 3), (128, 0, 128, 3), (128, 0, 256, 1), (64, 0, 128, 1), (32, 0, 16, 1), (32, 0, 4, 1)]
    res_of = {16: 1, 32: 1, 64: 2, 128: 4, 256: 4}                       # spatial divisor of the level a layer lives on
    for size, imgs_per_frame in ((80, 1), (64, 30)):                    # CenterNet 80x80, P2PNet 64x64 x 3 planes x 10 people
        for batch in (1, 2, 8, 32):
            for cin, cin2, cout, k in layers:
                d = res_of[max(cin, 16)]
                p = _conv_plan(lib, batch * imgs_per_frame, size // d, size // d, cin, cin2, cout, k)
                assert p["n_tile"] > 0 and p["n_tile"] * p["n_tiles"] >= cout
                assert p["smem"] <= 227 * 1024 - 1280                   # opt-in limit minus the static part
                assert p["tmem"] in (32, 64, 128, 256, 512) and p["tmem"] >= 4 * min(p["n_tile"], 32)
                assert 1 <= p["grid"] <= 148 * (2 if p["occ2"] else 1) and p["grid"] <= p["items"]
                if p["occ2"]:
                    assert 2 * (p["smem"] + 1280 + 1024) <= 228 * 1024 and p["tmem"] <= 256
                if p["resident"] == 2:
                    assert p["n_tiles"] > 1 and p["grid"] % p["n_tiles"] == 0 and p["grid"] >= p["n_tiles"]
                if p["resident"] == 0:
                    assert p["b_stages"] >= 3
    # the decisions documented in DESIGN.md 4.1
    assert _conv_plan(lib, 960, 64, 64, 32, 0, 32, 3)["occ2"] == 1 and _conv_plan(lib, 960, 64, 64, 16, 0, 16, 7)["occ2"] == 1
    assert _conv_plan(lib, 960, 64, 64, 16, 0, 32, 3)["occ2"] == 0      # loader-bound 16-channel 3x3: one CTA per SM
    assert _conv_plan(lib, 1, 20, 20, 128, 0, 128, 3) | {"smem": 0} == {"n_tile": 32, "n_tiles": 4, "resident": 2, "a_stages": 2,
                                                                         "b_stages": 1, "smem": 0, "occ2": 0, "grid": 24, "items": 24,
                                                                         "tmem": 128}
    big = _conv_plan(lib, 960, 16, 16, 128, 0, 128, 3)
    assert (big["n_tile"], big["resident"], big["tmem"]) == (128, 0, 512)           # streamed, full-width N at large batch
    assert _conv_plan(lib, 960, 32, 32, 64, 32, 64, 3)["resident"] == 1             # 19 weight blocks + 2 halo stages fit in 227 KB
    p30 = _conv_plan(lib, 30, 16, 16, 128, 64, 128, 3)                              # P2PNet at batch 1: 64-column tiles, weights
    assert (p30["n_tile"], p30["resident"], p30["grid"]) == (64, 0, 120)            # streamed row-wise (measured faster, round 2)
    # the 7x7 front conv of the J = 17 datasets (JP = 20 channels): two 16-channel K-blocks, 98 resident weight blocks
    j17 = _conv_plan(lib, 30, 64, 64, 20, 0, 16, 7)
    assert (j17["n_tile"], j17["resident"], j17["a_stages"]) == (16, 1, 2) and j17["smem"] <= 227 * 1024 - 1280
    # layers the tensor-core engine leaves to the CUDA-core kernel
    assert _conv_plan(lib, 1, 64, 64, 48, 0, 32, 7)["n_tile"] == 0 and _conv_plan(lib, 1, 16, 16, 128, 0, 512, 1)["n_tile"] == 0


@pytest.mark.parametrize("cin,cin2,cout,k,variant,cb", [(64, 32, 64, 3, 0, 32), (128, 0, 128, 3, 1, 32), (32, 16, 32, 3, 0, 32),
                                                        (16, 0, 16, 7, 0, 16), (16, 0, 32, 3, 0, 16), (128, 0, 256, 1, 2, 32),
                                                        (32, 0, 15, 1, 0, 32), (256, 0, 512, 1, 0, 32)])   # last: a backbone-sized layer (N2)
def test_tensor_core_weight_image_layout(built_library, cin, cin2, cout, k, variant, cb):
    """The host packer's shared-memory image against an independent statement of the layout k_conv_tc's descriptors read
    (include/fvp_b200.h, fvp_debug_pack_tc16), and the hi/lo split against NumPy float16 arithmetic."""
    import ctypes as C
    from fvp import capi
    lib = capi.load()
    rng = np.random.default_rng(cin * 7 + cout + k)
    ru = lambda a, b: -(-a // b) * b
    coutp, cinp, cin2p = ru(cout, 4), ru(cin, 16), (ru(cin2, 16) if cin2 else 0)
    rows = k * k * cinp + cin2p
    w = np.zeros((rows, coutp), np.float32)
    for tap in range(k * k):
        w[tap * cinp: tap * cinp + cin, :cout] = rng.normal(0, 0.2, (cin, cout)) * 10.0 ** rng.integers(-3, 1, (cin, cout))
    if cin2:
        w[k * k * cinp: k * k * cinp + cin2, :cout] = rng.normal(0, 0.2, (cin2, cout))
    n = C.c_longlong(0)
    assert lib.fvp_debug_pack_tc16(w.ctypes.data, cin, cin2, coutp, k, variant, cb, None, 0, C.byref(n)) == -2
    img = np.zeros(n.value, np.uint16)
    assert lib.fvp_debug_pack_tc16(w.ctypes.data, cin, cin2, coutp, k, variant, cb, img.ctypes.data, n.value, C.byref(n)) == 0
    # independent reconstruction of the contract
    npad = ru(coutp, 16)
    cap = {0: 128, 1: 32, 2: 64}[variant]
    n_tile = npad if npad <= cap else cap
    n_tiles = -(-npad // n_tile)
    chunks, shift = cb // 8, (1 if cb == 32 else 2)
    hi = w.astype(np.float16)
    lo = ((w - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    expect = np.zeros(n.value, np.uint16)
    pos = 0
    for taps, cp, rowbase in ((k * k, cinp, 0),) + (((1, cin2p, k * k * cinp),) if cin2 else ()):
        for c0 in range(0, ru(cp, cb), cb):
            for tap in range(taps):
                for nt in range(n_tiles):
                    for part in (hi, lo):
                        blk = np.zeros((n_tile, cb), np.uint16)
                        for r in range(n_tile):
                            col = nt * n_tile + r
                            for cc in range(cb):
                                ci = c0 + cc
                                v = part[rowbase + tap * cp + ci, col] if (col < coutp and ci < cp) else np.float16(0)
                                chunk = (cc >> 3) ^ ((r >> shift) & (chunks - 1))
                                blk[r, chunk * 8 + (cc & 7)] = np.float16(v).view(np.uint16)
                        expect[pos: pos + blk.size] = blk.ravel()
                        pos += blk.size
    assert pos == n.value
    assert np.array_equal(img, expect)
    # the pair carries 22 significant bits of every weight
    back = hi.astype(np.float64) + lo.astype(np.float64) / 2048.0
    nz = np.abs(w) > 1e-4                      # below fp16's normal range of the lo term the split degrades gracefully
    assert np.max(np.abs(back - w)[nz] / np.abs(w)[nz]) < 2.0 ** -20


def test_two_cta_conv_variants_do_not_spill_more_than_measured(built_library):
    """The 72-register variants of k_conv_tc (two CTAs per SM) run beside 2 x 111 KB of shared memory, i.e. with 28 KB of L1:
    ptxas spills beyond the sizes measured fast on B200 (profiles/r01_s5_conv_layers_final.txt) cost up to 45 % on the
    full-resolution layers.  The build log is the evidence (csrc/build.sh keeps -Xptxas -v output per source file)."""
    import re
    log = open(os.path.join(os.path.dirname(built_library), "csrc", "build", "fvp_conv_tc.log")).read()
    spills = {}
    for m in re.finditer(r"k_conv_tcILi(\d)ELi(\d)ELb(\d)EE.*?\n.*?(\d+) bytes spill stores", log):
        spills[(int(m.group(1)), int(m.group(2)), int(m.group(3)))] = int(m.group(4))
    legacy = {(m, o, 0) for m in (0, 1, 2) for o in (1, 2)}              # fp32 activations staged by loader warps
    tma = {(m, o, 1) for m in (1, 2) for o in (1, 2)}                    # split activations fetched by TMA tensor loads
    assert set(spills) == legacy | tma, spills
    assert spills[(1, 2, 0)] <= 32 and spills[(2, 2, 0)] <= 128 and spills[(1, 1, 0)] <= 32, spills
    assert spills[(1, 2, 1)] <= 32 and spills[(1, 1, 1)] <= 32 and spills[(2, 2, 1)] <= 64, spills
    # the split-activation path is real TMA: tensor-map loads (UTMALDG) next to the bulk weight copies (UBLKCP)
    import shutil
    import subprocess
    if shutil.which("cuobjdump"):
        sass = subprocess.run(["cuobjdump", "-sass", built_library], capture_output=True, text=True).stdout
        assert sass.count("UTMALDG") >= 8 and "UBLKCP" in sass and "UTCHMMA" in sass


def test_cluster_proposal_kernel_exchanges_through_distributed_shared_memory(built_library):
    """k_proposals_cluster (C2CNet on an 8-CTA cluster per column): the layer outputs travel as 16-byte st.async stores
    (SASS STAS.128) onto per-layer mbarriers, the weight slices arrive by bulk copies (UBLKCP), and the only cluster barriers
    are the one after the mbarrier initialisation and the one before exit - none inside the layer loop."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "k_proposals_cluster", built_library], capture_output=True, text=True).stdout
    if "STAS" not in sass:                      # older cuobjdump: -fun wants the mangled name; fall back to the whole file
        full = subprocess.run(["cuobjdump", "-sass", built_library], capture_output=True, text=True).stdout
        start = full.index("k_proposals_cluster")
        nxt = full.find("Function :", start)
        sass = full[start:nxt if nxt > 0 else len(full)]
    assert sass.count("STAS.128") >= 19 and "STAS.64" in sass          # every layer sends vectors; the first layer 8-byte pairs
    assert sass.count("UBLKCP") >= 1
    assert sass.count("UCGABAR_ARV") == 2 and sass.count("UCGABAR_WAIT") == 2


def test_k3_ships_the_three_shuffle_tap_exchange(built_library):
    """The per-person back-projection kernel of 64-byte records (J = 13..16) hands a depth's taps to its lane group with
    three shuffles (offset + two fractions) per depth and view, 12 per unrolled view iteration - the form measured 3 %
    faster than five (profiles/r02_k3_xch.txt); the J = 17 instantiation keeps the five-shuffle form (20)."""
    import re
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    full = subprocess.run(["cuobjdump", "-sass", built_library], capture_output=True, text=True).stdout
    counts = {}
    for m in re.finditer(r"Function : (\S*k3_jln_patch\S*)", full):
        end = full.find("Function :", m.end())
        counts[m.group(1)] = full[m.end():end if end > 0 else len(full)].count("SHFL.IDX")
    by = lambda tag: [n for k, n in counts.items() if tag in k]
    assert by("ILi4ELi4ELi64ELi1E") == [12]                      # <CG 4, 4 CTAs/SM, 64-byte pixel stride, XCH 1>: shipped for JG == 4
    assert by("ILi8ELi2ELi0ELi0E") == [40]                       # <CG 8, XCH 0>: 5 shuffles x 8 depths
    assert not by("ILi4ELi4ELi64ELi0E")                          # the five-shuffle 64-byte variant is no longer instantiated


# ---- reference-facing module -----------------------------------------------------------------------
def test_model_has_reference_state_dict_and_no_cpu_path(built_library, golden):
    import models
    g = golden("panoptic_none_valid")
    cfg = g.cfg
    cfg.DEVICE = "cpu"
    m = models.faster_voxelpose.get(cfg)
    keys = list(m.state_dict().keys())
    assert keys == [k for k, _, _ in netspec.param_table(15)]
    res = m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in g.weights.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert hasattr(m, "pose_net") and hasattr(m, "joint_net") and hasattr(m.pose_net, "center_net")
    m.eval()
    with pytest.raises(RuntimeError):        # product path refuses to run without the GPU: no fallback
        m(meta={"seq": ["s"]}, input_heatmaps=torch.from_numpy(g.heatmaps), cameras={"s": g.cams},
          resize_transform=torch.as_tensor(g.resize, dtype=torch.float))
    m.train()
    with pytest.raises(NotImplementedError):
        m(meta={"seq": ["s"]}, input_heatmaps=torch.from_numpy(g.heatmaps), cameras={"s": g.cams},
          resize_transform=torch.as_tensor(g.resize, dtype=torch.float))


def test_weight_change_tag_is_cheap_and_sees_every_kind_of_update(built_library, golden):
    """The plugin re-packs device weights when a parameter changed; the per-forward check must notice in-place edits,
    load_state_dict (copy and assign) and module conversions, and must not walk state_dict() every call."""
    import time
    import models
    g = golden("panoptic_none_valid")
    cfg = g.cfg
    cfg.DEVICE = "cpu"
    m = models.faster_voxelpose.get(cfg).eval()
    t0 = m._tag()
    assert m._tag() == t0 and len(t0) == 485
    with torch.no_grad():
        next(m.parameters()).mul_(1.0)                                   # in-place edit: version counter
    t1 = m._tag()
    assert t1 != t0
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in g.weights.items()}, strict=False)
    t2 = m._tag()
    assert t2 != t1
    m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()}, assign=True)       # tensors replaced
    t3 = m._tag()
    assert t3 != t2 and len(t3) == 485
    m.double()                                                           # _apply: storage replaced
    assert m._tag() != t3
    m.float()
    def best_ms(fn, n=20, reps=5):                                       # best of several repetitions: robust on a busy host
        best = float("inf")
        for _ in range(reps):
            t = time.perf_counter()
            for _ in range(n):
                fn()
            best = min(best, (time.perf_counter() - t) / n * 1e3)
        return best
    per_call_ms = best_ms(m._tag)
    walk_ms = best_ms(lambda: m.state_dict(keep_vars=True))
    assert per_call_ms < 0.7 * walk_ms, (per_call_ms, walk_ms)          # measured ~0.15 x; the bound only guards the design


def test_product_code_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)


# ---- frame sharding ----------------------------------------------------------------------------------
def test_shard_bounds_cover_all_frames():
    for n in (1, 7, 32, 256):
        for w in (1, 2, 3, 8):
            spans = [fdist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from fvp import dist as fdist
rank, world, _ = fdist.init_from_env("gloo")
N = int(sys.argv[2])
full = torch.arange(N * 10 * 15 * 5, dtype=torch.float32).view(N, 10, 15, 5) * 0.5      # stands for fused_poses
out = fdist.sharded_forward(lambda lo, hi: full[lo:hi].clone(), N, rank, world)
assert torch.equal(out, full), (rank, out.shape)
# the bench's form: equal shards of S rows, ONE all_gather_into_tensor into one contiguous rank-major buffer
S = 4
mine = full[:S] + 1000.0 * rank
buf = torch.empty((world * S, 10, 15, 5))
got = fdist.gather_shards(mine, buf)
assert got.data_ptr() == buf.data_ptr() and got.is_contiguous()
assert all(torch.equal(got[r * S:(r + 1) * S], full[:S] + 1000.0 * r) for r in range(world))
cores = fdist.pin_rank_to_cores(rank, world)
assert len(cores) >= 1 and set(cores) == set(os.sched_getaffinity(0))
if rank == 0: print("GATHER_OK", N, world)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("nframes", [8, 7])
def test_frame_sharded_gather_world2_gloo(tmp_path, nframes):
    """N>1 path: 2 ranks (gloo, CPU) shard frames contiguously and one all_gather restores frame order,
    bit-identical to the single-rank tensor, also for ragged shards."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29600 + nframes))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29700 + nframes), str(script), PKG,
                        str(nframes)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _fake_sysfs(root, node_cpus, siblings, gpus):
    """node_cpus {node: 'cpulist'}, siblings {cpu: 'cpulist'}, gpus {pci id: node} -> a sysfs-shaped tree under root."""
    for node, cl in node_cpus.items():
        d = root / "devices/system/node" / ("node%d" % node)
        d.mkdir(parents=True)
        (d / "cpulist").write_text(cl + "\n")
    for cpu, cl in siblings.items():
        d = root / "devices/system/cpu" / ("cpu%d" % cpu) / "topology"
        d.mkdir(parents=True)
        (d / "thread_siblings_list").write_text(cl + "\n")
    for pci, node in gpus.items():
        d = root / "bus/pci/devices" / pci
        d.mkdir(parents=True)
        (d / "numa_node").write_text("%d\n" % node)
    return str(root)


def test_rank_core_plan_is_numa_aware_and_keeps_hyperthreads_together(tmp_path):
    """Two sockets of 8 physical cores with hyper-threads numbered after the first threads (0-7,16-23 | 8-15,24-31), eight
    GPUs, four per socket: every rank gets two whole physical cores of its GPU's socket; unknown topology, a lopsided
    cpuset or too few cores fall back to the even split / no pinning."""
    from fvp import dist as fdist
    sib = {c: "%d,%d" % (c % 16, c % 16 + 16) for c in range(32)}
    gpus = {"0000:%02x:00.0" % (0x10 + i): (0 if i < 4 else 1) for i in range(8)}
    sysfs = _fake_sysfs(tmp_path, {0: "0-7,16-23", 1: "8-15,24-31"}, sib, gpus)
    assert fdist._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert fdist.gpu_numa_node("0000:14:00.0", sysfs) == 1 and fdist.gpu_numa_node("0000:99:00.0", sysfs) is None
    nodes = [0, 0, 0, 0, 1, 1, 1, 1]
    plan = fdist.plan_rank_cores(list(range(32)), 8, nodes, sysfs)
    assert plan[0] == [0, 1, 16, 17] and plan[3] == [6, 7, 22, 23] and plan[4] == [8, 9, 24, 25] and plan[7] == [14, 15, 30, 31]
    flat = [c for p in plan for c in p]
    assert sorted(flat) == list(range(32))                                   # disjoint, complete
    # GPUs enumerated socket 1 first: the plan follows the GPUs, not the rank order
    plan = fdist.plan_rank_cores(list(range(32)), 8, nodes[::-1], sysfs)
    assert plan[0] == [8, 9, 24, 25] and plan[7] == [6, 7, 22, 23]
    # unknown node of one GPU -> even split of physical cores in rank order (siblings still together)
    plan = fdist.plan_rank_cores(list(range(32)), 8, [0, 0, 0, None, 1, 1, 1, 1], sysfs)
    assert plan[0] == [0, 1, 16, 17] and plan[4] == [8, 9, 24, 25]
    # a cpuset that leaves socket 1 two cores for four ranks -> even split over everything allowed
    allowed = list(range(8)) + [8, 9] + list(range(16, 24)) + [24, 25]
    plan = fdist.plan_rank_cores(allowed, 8, nodes, sysfs)
    assert plan[0] == [0, 16] and plan[7] == [7, 23]                         # 10 physical cores // 8 ranks = 1 core each
    assert len({c for p in plan for c in p}) == sum(len(p) for p in plan) and {c for p in plan for c in p} <= set(allowed)
    # fewer physical cores than ranks -> nobody is pinned
    assert fdist.plan_rank_cores([0, 16], 2, [0, 0], sysfs) == [[], []]
    # no topology files at all (a bare container): every CPU is its own core, even split
    assert fdist.plan_rank_cores(list(range(8)), 2, [None, None], str(tmp_path / "nothing")) == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_rank_core_plan_does_not_trade_cores_for_locality(tmp_path):
    """All eight GPUs report node 0 of a two-node box: sharing node 0's 8 physical cores would halve every rank's
    share, so the even split over both nodes stays."""
    from fvp import dist as fdist
    sib = {c: "%d,%d" % (c % 16, c % 16 + 16) for c in range(32)}
    sysfs = _fake_sysfs(tmp_path, {0: "0-7,16-23", 1: "8-15,24-31"}, sib, {})
    plan = fdist.plan_rank_cores(list(range(32)), 8, [0] * 8, sysfs)
    assert plan[0] == [0, 1, 16, 17] and plan[7] == [14, 15, 30, 31]


def test_gpu_numa_nodes_from_device_properties(tmp_path, monkeypatch):
    """PCI address formatting of torch's device properties -> sysfs lookup; a failing query reads as 'unknown'."""
    import types
    from fvp import dist as fdist
    sysfs = _fake_sysfs(tmp_path, {0: "0-3", 1: "4-7"}, {}, {"0000:1b:00.0": 0, "0001:e3:00.0": 1, "0000:2a:00.0": -1})
    props = [types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1B, pci_device_id=0),
             types.SimpleNamespace(pci_domain_id=1, pci_bus_id=0xE3, pci_device_id=0),
             types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x2A, pci_device_id=0)]

    def fake_props(i):
        if i >= len(props):
            raise RuntimeError("invalid device ordinal")
        return props[i]
    monkeypatch.setattr(torch.cuda, "get_device_properties", fake_props)
    assert fdist.local_gpu_numa_nodes(4, sysfs) == [0, 1, None, None]
    import json
    json.dumps(fdist.local_gpu_numa_nodes(4, sysfs))                         # goes into the bench line as is


def test_pinning_never_raises_on_odd_sysfs(tmp_path):
    """Pinning is an optimisation: malformed topology files must not take a multi-GPU run down."""
    from fvp import dist as fdist
    root = tmp_path
    d = root / "devices/system/node/node0"
    d.mkdir(parents=True)
    (d / "cpulist").write_text("0-,x\n")                                     # garbage
    before = os.sched_getaffinity(0)
    assert fdist.pin_rank_to_cores(0, 2, [0, 0], str(root)) == []
    assert os.sched_getaffinity(0) == before
