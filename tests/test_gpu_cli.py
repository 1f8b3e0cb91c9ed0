"""GPU end-to-end: the generate and evaluate entry points on a small synthetic corpus, against the
oracle pipeline run on the same WAV files (BASELINE configs[0] in miniature)."""
import os

import numpy as np
import pytest
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def corpus(tmp_path_factory):
    from nafp_b200 import synth
    root = tmp_path_factory.mktemp("corpus")
    src = root / "music"
    for sub in ("test-dummy-db-100k-full/a", "test-query-db-500-30s/query/a", "test-query-db-500-30s/db/a"):
        os.makedirs(src / sub)
    for i in range(12):                                    # dummy: 12 x 12 s
        synth.write_wav(str(src / "test-dummy-db-100k-full/a" / f"d{i:03d}.wav"), synth.synth_track(100 + i, 96000))
    for i in range(4):                                     # db / query: 4 x 10 s, query = db + noise at 5 dB SNR
        x = synth.synth_track(200 + i, 80000)
        synth.write_wav(str(src / "test-query-db-500-30s/db/a" / f"s{i:03d}.wav"), x)
        synth.write_wav(str(src / "test-query-db-500-30s/query/a" / f"s{i:03d}.wav"), synth.add_noise_snr(x, 5.0, i))
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "default.yaml")))
    cfg['DIR']['SOURCE_ROOT_DIR'] = str(src) + "/"
    cfg['DIR']['LOG_ROOT_DIR'] = str(root / "logs") + "/"
    cfg['DIR']['OUTPUT_ROOT_DIR'] = str(root / "logs" / "emb") + "/"
    cfg['DATA_SEL']['TEST_DUMMY_DB'] = '100k_full_icassp'
    cfg['BSZ']['TS_BATCH_SZ'] = 25
    return cfg, src


def test_generate_then_evaluate(corpus):
    import glob
    from nafp_b200.eval.eval_search import run_eval
    from nafp_b200.model import weights as W
    from nafp_b200.model.generate import generate_fingerprint
    from oracle import fingerprinter as ofp
    from oracle import melspec, segments, seq_match
    from oracle.flat_index import FlatL2
    cfg, src = corpus
    generate_fingerprint(cfg, 'random-init:7', None, None, None, False)
    emb_dir = cfg['DIR']['OUTPUT_ROOT_DIR'] + '/random-init:7/0/'
    w = W.init_weights(7)
    got = {}
    for key, sub in (("dummy_db", "test-dummy-db-100k-full/"), ("query", "test-query-db-500-30s/query/"),
                     ("db", "test-query-db-500-30s/db/")):
        shape = np.load(emb_dir + f"{key}_shape.npy")
        arr = np.memmap(emb_dir + f"{key}.mm", dtype='float32', mode='r', shape=tuple(shape))
        files = sorted(glob.glob(os.path.join(str(src), sub) + '**/*.wav', recursive=True))
        ref = np.concatenate([ofp.fingerprinter(melspec.melspec_layer(b), w) for b in segments.batches(files, bsz=25)])
        assert arr.shape == ref.shape == (shape[0], 128)
        cos_min, err_max = (np.asarray(arr) * ref).sum(1).min(), np.abs(np.asarray(arr) - ref).max()
        print(f"{key}: {shape[0]} fingerprints, min cosine {cos_min:.7f}, max abs err {err_max:.3e}")
        assert cos_min >= 0.9999 and err_max <= 1e-3
        got[key] = np.array(arr)
    assert got["dummy_db"].shape[0] == 12 * 23 and got["db"].shape[0] == 4 * 19

    ids_path = emb_dir + "ids.npy"
    test_ids = np.arange(0, got["query"].shape[0] - 9, 3)
    np.save(ids_path, test_ids)
    rates = run_eval(emb_dir, index_type='l2', test_ids=ids_path, test_seq_len='1 3 5 9', k_probe=20, display_interval=5,
                     live=False)
    raw = np.load(emb_dir + "raw_score.npy")
    assert (np.load(emb_dir + "test_ids.npy") == test_ids).all()
    # the oracle evaluation on the SAME (GPU-generated) fingerprints must agree flag by flag
    idx = FlatL2(128)
    idx.add(got["dummy_db"])
    idx.add(got["db"])
    raw_o, _ = seq_match.evaluate(idx, got["query"], np.concatenate([got["dummy_db"], got["db"]]), len(got["dummy_db"]),
                                  test_ids, [1, 3, 5, 9], 20)
    assert raw.shape == raw_o.shape == (len(test_ids), 16)
    assert (raw != raw_o).sum() <= 1                      # at most one fp32-tie flip
    assert np.abs(np.array(rates) - seq_match.hit_rates(raw_o, 4)).max() <= 100.0 / len(test_ids) + 1e-9
    assert not os.path.exists(emb_dir + "dummy_db.mm.bak")
    assert os.path.getsize(emb_dir + "dummy_db.mm") == 12 * 23 * 128 * 4      # inputs are never extended on disk


def _run_generate_ranks(cfg_path, out_dir, world, tmp):
    """`run.py generate` once per rank, all ranks on GPU 0 (NAFP_DEVICE), the way torchrun would start them."""
    import socket
    import subprocess
    import sys
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), NAFP_DEVICE="0",
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PYTHONPATH=ROOT)
        log = open(os.path.join(tmp, f"gen_w{world}_r{rank}.log"), "w")
        procs.append((subprocess.Popen([sys.executable, "-m", "nafp_b200.run", "generate", "random-init:7", "-c", cfg_path,
                                        "-o", out_dir], env=env, cwd=ROOT, stdout=log, stderr=subprocess.STDOUT), log))
    for p, log in procs:
        rc = p.wait(timeout=600)
        log.close()
        assert rc == 0, open(log.name).read()[-2000:]


def test_generate_sharded_over_two_ranks_is_byte_identical(corpus, tmp_path):
    """SURVEY §8 e / north_star: generation is split by segment batch with no collective.  Two ranks (started like
    torchrun starts them, gloo barrier around the file creation, both on this box's GPU) must write exactly the
    bytes one rank writes -- batches are whole TS_BATCH_SZ groups, so every group's log-mel max is unchanged."""
    cfg, _ = corpus
    cfg_path = str(tmp_path / "cfg.yaml")
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f)
    one, two = str(tmp_path / "w1"), str(tmp_path / "w2")
    _run_generate_ranks(cfg_path, one, 1, str(tmp_path))
    _run_generate_ranks(cfg_path, two, 2, str(tmp_path))
    for key, rows in (("dummy_db", 12 * 23), ("query", 4 * 19), ("db", 4 * 19)):
        a = open(f"{one}/random-init:7/0/{key}.mm", "rb").read()
        b = open(f"{two}/random-init:7/0/{key}.mm", "rb").read()
        assert len(a) == rows * 128 * 4
        assert a == b, f"{key}.mm differs between the 1-rank and the 2-rank run"
        assert (np.load(f"{one}/random-init:7/0/{key}_shape.npy") == np.load(f"{two}/random-init:7/0/{key}_shape.npy")).all()
    emb = np.frombuffer(open(f"{two}/random-init:7/0/db.mm", "rb").read(), np.float32).reshape(-1, 128)
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-5)          # no rank left its rows unwritten
