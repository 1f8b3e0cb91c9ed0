"""Host logic and the C-ABI surface, no GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from nafp_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "nafp.h")).read()
    declared = sorted(set(re.findall(r"\b(nafp_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 45
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/nafp.h but not exported by libnafp.so"
        assert name in _lib.EXPORTS, f"{name} has no ctypes signature in _lib.py"
    assert _lib.lib.nafp_version() >= 100


def test_no_gpu_fails_loudly():
    from nafp_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.NafpError):
        _lib.Context(0)
    from nafp_b200.eval.utils.get_index import get_index
    with pytest.raises(_lib.NafpError):
        get_index('l2', np.zeros((4, 128), np.float32), (4, 128), use_gpu=False)     # no CPU path by design


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "neural-audio-fp_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(".py") or fn.endswith(".cu") or fn.endswith(".h") or fn.endswith(".cuh"):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{fn} imports oracle/"


def test_shard_bounds_cover_and_halo():
    from nafp_b200.dist import shard_bounds, shard_with_halo
    for n, w in ((56_029_500, 8), (1000, 3), (7, 8), (29500, 2)):
        edges = [shard_bounds(n, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 1
    assert shard_with_halo(1000, 0, 2, 18) == (0, 500, 518)
    assert shard_with_halo(1000, 1, 2, 18) == (500, 1000, 1000)
    from nafp_b200.model.generate import _shard
    got = [_shard(95, r, 4) for r in range(4)]
    assert got[0][0] == 0 and got[-1][1] == 95 and all(got[i][1] == got[i + 1][0] for i in range(3))


def test_config_guards_and_cli_surface(tmp_path):
    import yaml
    from click.testing import CliRunner
    from nafp_b200.model import fp
    from nafp_b200.run import cli
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "default.yaml")))
    fp._check_model_cfg(cfg)
    bad = yaml.safe_load(open(os.path.join(ROOT, "config", "default.yaml")))
    bad['MODEL']['N_MELS'] = 128
    with pytest.raises(NotImplementedError):
        fp._check_model_cfg(bad)
    r = CliRunner().invoke(cli, ["generate", "--help"])
    assert r.exit_code == 0 and "--skip_dummy" in r.output and "--source" in r.output
    from nafp_b200.eval.eval_search import eval_faiss
    r = CliRunner().invoke(eval_faiss, ["--help"])
    assert r.exit_code == 0
    for opt in ("--emb_dummy_dir", "--index_type", "--nogpu", "--max_train", "--test_seq_len", "--test_ids", "--k_probe",
                "--display_interval"):
        assert opt in r.output


def test_eval_helpers(tmp_path):
    from nafp_b200.eval import eval_search as es
    from oracle import seq_match
    a = np.random.default_rng(0).standard_normal((10, 128)).astype(np.float32)
    d = str(tmp_path) + "/"
    m = np.memmap(d + "db.mm", dtype='float32', mode='w+', shape=a.shape)
    m[:] = a
    m.flush()
    np.save(d + "db_shape.npy", a.shape)
    data, shape = es.load_memmap_data(d, "db", display=False)
    assert tuple(shape) == (10, 128) and (np.asarray(data) == a).all()
    assert tuple(es.load_memmap_data(d, "db", shape_only=True)) == (10, 128)
    ids = es.select_test_ids('icassp', 29500, [1, 19])
    assert len(ids) == 2000 and ids.max() == 29492
    assert (es.select_test_ids('all', 100, [1, 19]) == np.arange(81)).all()
    assert len(es.select_test_ids('7', 100, [1, 3], rng=np.random.default_rng(0))) == 7
    for pred, gt in ((np.array([7, 3, 9, -1, -1]), 8), (np.array([7, 8, 9]), 8), (np.array([-1, -1]), 3)):
        assert es.hit_flags(pred, gt) == seq_match.hit_flags(pred[pred >= 0], gt)


def test_print_table_summary_format():
    from nafp_b200.eval.utils.print_table import PrintTable
    pt = PrintTable([1, 3, 19], ['Top1 exact', 'Top1 near', 'Top3 exact', 'Top10 exact'], live=False)
    pt.update_table(([50.0, 90.0, 100.0],) * 4)
    pt.update_counter(9, 10, 1.234)
    lines = pt.summary_lines()
    assert lines[0] == '========= Top1 hit rate (%) of segment-level search ========='
    assert '(1s)' in lines[3] and '(2s)' in lines[3] and '(10s)' in lines[3]
    assert lines[-1] == 'average search + evaluation time 1.23 ms/query'
    assert lines[5].startswith('  Top1 exact  ') and '50.00' in lines[5]


def test_weights_roundtrip(tmp_path):
    from nafp_b200.model import weights
    w = weights.init_weights(3, randomize_affine=True)
    p = str(tmp_path / "ckpt-1.npz")
    weights.save_weights(p, w)
    w2 = weights.load_weights(p)
    assert set(w) == set(w2) and all((w[k] == w2[k]).all() for k in w)
    assert w["conv0_a_w"].shape == (1, 3, 1, 128) and w["conv7_b_w"].shape == (3, 1, 1024, 1024)
    assert w["ln0_a_g"].shape == (256, 16, 128) and w["div_w1"].shape == (128, 8, 32)
    # glorot_uniform bound of Keras for conv1_b: sqrt(6 / (3*128 + 3*128))
    assert np.abs(w["conv1_b_w"]).max() <= np.sqrt(6 / 768) + 1e-7


def test_generate_reads_the_launcher_environment(monkeypatch):
    """run.py generate under torchrun: rank / world / GPU come from RANK, WORLD_SIZE, LOCAL_RANK (NAFP_DEVICE overrides)."""
    from nafp_b200.model import generate as G
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "NAFP_DEVICE"):
        monkeypatch.delenv(k, raising=False)
    assert G.distributed_env() == (0, 1, 0)
    monkeypatch.setenv("RANK", "3"); monkeypatch.setenv("WORLD_SIZE", "8"); monkeypatch.setenv("LOCAL_RANK", "3")
    assert G.distributed_env() == (3, 8, 3)
    monkeypatch.setenv("NAFP_DEVICE", "0")
    assert G.distributed_env() == (3, 8, 0)
    G._barrier_fn(1)()                                  # single rank: no process group is touched
    with pytest.raises(ValueError):
        G.generate_fingerprint({}, 'random-init', None, None, None, True, rank=2, world_size=2)


def test_cli_limits_are_checked_up_front():
    """limits of the matcher / IVF kernels surface as a clear ValueError before any data is touched."""
    from nafp_b200.eval.eval_search import _check_limits
    _check_limits('l2', [1, 3, 5, 9, 11, 19], 20)
    _check_limits('ivfpq', [1, 19], 51)
    _check_limits('ivfpq-rr', [1, 19], 20)
    for args in (('l2', [1, 33], 20), ('l2', [19], 60), ('ivfpq', [1], 110), ('ivfpq-rr', [1], 40), ('l2', [], 20)):
        with pytest.raises(ValueError):
            _check_limits(*args)


def test_bench_reads_roofline_traffic_from_the_committed_profiles():
    """bench.py's `roofline.traffic` is parsed from the ncu summaries under profiles/, not typed in."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    scan, f1 = bench.ncu_traffic("flat_scan_kernel", "r*_prof_scan*summary.csv")
    assert f1 and f1.startswith("profiles/r2_") and 14.3e9 < scan < 15.5e9          # 56,029,500 rows x 256 B = 14.34 GB algorithmic
    mel, f2 = bench.ncu_traffic("logmel_kernel", "r*_prof_logmel*summary.csv")
    assert f2 and 1.0e8 < mel < 4.0e8                                                # 4,000 segments x 64,768 B = 0.26 GB algorithmic
    assert bench.ncu_traffic("no_such_kernel", "r*_prof_scan*summary.csv") == (None, None)
