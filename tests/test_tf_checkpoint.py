"""TensorFlow-checkpoint reader against an independent writer of the same published format
(tests/util_tf_bundle.py).  PARITY UNPINNED against TensorFlow itself: none is installed and the reference ships
no checkpoint; what IS pinned: the CRC-32C known answer, the table magic, and the parameter count of
``model/fp/nnfp.py:270-274``."""
import os

import numpy as np
import pytest

from nafp_b200.model import arch, tf_checkpoint as T
from nafp_b200.model.weights import init_weights
from util_tf_bundle import reference_variable_names, write_bundle


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                    # CRC-32C check value
    assert T.crc32c(b"") == 0
    assert T.crc32c(b"56789", T.crc32c(b"1234")) == 0xE3069283     # continuation
    x = np.random.default_rng(0).integers(0, 256, 300001, dtype=np.uint8)
    assert T._crc32c_fast(x) == T.crc32c(x.tobytes())              # lane-parallel form = byte-serial form
    assert T.crc32c_mask(0) == 0xA282EAD8


def test_snappy_decoder_on_a_hand_made_stream():
    # "abcabcabcabcX": literal 'abc', copy(offset 3, length 9) as a 2-byte-offset copy, literal 'X'
    stream = bytes([13]) + bytes([2 << 2]) + b"abc" + bytes([(9 - 1) << 2 | 2, 3, 0]) + bytes([0 << 2]) + b"X"
    assert T._snappy_decompress(stream) == b"abcabcabcabcX"


def test_full_checkpoint_round_trip(tmp_path):
    w = init_weights(3, randomize_affine=True)
    tensors = reference_variable_names(w)
    assert sum(v.size for v in tensors.values()) == arch.n_params() == 16_939_008       # nnfp.py:270-274 at (256,32,1)
    # what a training checkpoint holds besides the model: counters, optimizer slots, the object graph
    tensors["save_counter/.ATTRIBUTES/VARIABLE_VALUE"] = np.array(7, np.int64)
    slot = "model/front_conv/layer_with_weights-0/conv2d_1x3/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"
    tensors[slot] = np.zeros((1, 3, 1, 128), np.float32)
    prefix = str(tmp_path / "ckpt-100")
    write_bundle(prefix, tensors, strings={"_CHECKPOINTABLE_OBJECT_GRAPH": b"\x0a\x02hi"})
    got = T.load_tf_checkpoint(prefix)
    assert set(got) == set(w)
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])
    everything = T.read_bundle(prefix)
    assert int(everything["save_counter/.ATTRIBUTES/VARIABLE_VALUE"].reshape(-1)[0]) == 7 and slot in everything
    assert T.latest_checkpoint(str(tmp_path)) == (100, prefix)
    with open(tmp_path / "checkpoint", "w") as f:
        f.write('model_checkpoint_path: "ckpt-100"\nall_model_checkpoint_paths: "ckpt-100"\n')
    assert T.latest_checkpoint(str(tmp_path)) == (100, prefix)
    npz = T.convert(prefix)
    from nafp_b200.model.weights import load_weights
    back = load_weights(npz)
    assert all((back[k] == w[k]).all() for k in w)


def test_variables_reached_through_the_forward_sequential_are_accepted(tmp_path):
    w = init_weights(4, randomize_affine=True)
    tensors = {}
    slots = {"conv2d_1x3": 0, "BN_1x3": 1, "conv2d_3x1": 2, "BN_3x1": 3}
    for k, v in reference_variable_names(w).items():
        for name, n in slots.items():
            k = k.replace(f"/{name}/", f"/forward/layer_with_weights-{n}/")
        tensors[k] = v
    prefix = str(tmp_path / "ckpt-1")
    write_bundle(prefix, tensors, block_size=300)               # many small blocks: exercises the index block
    got = T.load_tf_checkpoint(prefix)
    assert all((got[k] == w[k]).all() for k in w)


def test_corruption_and_mismatch_are_reported(tmp_path):
    w = init_weights(5)
    tensors = reference_variable_names(w)
    prefix = str(tmp_path / "ckpt-2")
    write_bundle(prefix, tensors)
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    raw[1000] ^= 0x40
    open(data, "wb").write(raw)
    with pytest.raises(ValueError, match="checksum"):
        T.load_tf_checkpoint(prefix)
    T.load_tf_checkpoint(prefix, verify=False)                  # explicit opt-out
    raw[1000] ^= 0x40
    open(data, "wb").write(raw)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[20] ^= 1
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError, match="checksum"):
        T.load_tf_checkpoint(prefix)
    # a model of another geometry
    bad = dict(tensors)
    k = "model/front_conv/layer_with_weights-3/conv2d_3x1/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    bad[k] = np.zeros((3, 1, 128, 256), np.float32)
    del bad["model/div_enc/split_fc_layers/5/layer_with_weights-1/bias/.ATTRIBUTES/VARIABLE_VALUE"]
    prefix2 = str(tmp_path / "ckpt-3")
    write_bundle(prefix2, bad)
    with pytest.raises(ValueError) as e:
        T.load_tf_checkpoint(prefix2)
    assert "expected (3, 1, 256, 256)" in str(e.value) and "missing conv3_b_w" in str(e.value) and "slice 5 b2" in str(e.value)
    with pytest.raises(ValueError, match="bad magic"):
        open(str(tmp_path / "x.index"), "wb").write(b"\0" * 100)
        T.read_index(str(tmp_path / "x.index"))
    assert not os.path.exists(str(tmp_path / "nothing.index"))
