"""TEST INFRASTRUCTURE: a minimal WRITER of TensorFlow's tensor-bundle checkpoint format (``ckpt-N.index`` SSTable +
``ckpt-N.data-00000-of-00001``), restated from the published format (tensorflow/core/util/tensor_bundle,
tensorflow/core/lib/io/table_builder, block_builder, format) independently of the reader under test: varint-coded
prefix-compressed entries with a restart point every 16 keys, 5-byte block trailers (compression byte + masked
crc32c), an index block of separator keys -> block handles, an empty metaindex block and the 48-byte footer."""
import struct

import numpy as np

from nafp_b200.model.tf_checkpoint import crc32c, crc32c_mask     # the checksum itself has a known-answer test

_DT = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}
DT_STRING = 7


def _vi(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(num, wt, payload):
    return _vi(num << 3 | wt) + payload


def _entry(dtype, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, _vi(len(d)) + d) for d in (_field(1, 0, _vi(s)) for s in shape))
    msg = _field(1, 0, _vi(dtype))
    msg += _field(2, 2, _vi(len(dims)) + dims)
    if offset:
        msg += _field(4, 0, _vi(offset))
    msg += _field(5, 0, _vi(size))
    msg += _field(6, 5, struct.pack("<I", crc))
    return msg


class _Block:
    def __init__(self, restart_interval=16):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _vi(shared) + _vi(len(key) - shared) + _vi(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def write_bundle(prefix, tensors, strings=None, block_size=4096):
    """``tensors``: {key: ndarray}; ``strings``: {key: bytes} stored as DT_STRING scalars (ignored by readers of
    numeric data).  Keys are written in sorted (bytewise) order, as TensorFlow does."""
    data = bytearray()
    entries = {b"": _field(1, 0, _vi(1)) + _field(3, 2, _vi(2) + _field(1, 0, _vi(1)))}     # header: num_shards 1, version
    for key in sorted(tensors):
        a = np.ascontiguousarray(tensors[key])
        raw = a.tobytes()
        entries[key.encode()] = _entry(_DT[a.dtype], a.shape, len(data), len(raw), crc32c_mask(_crc(raw)))
        data += raw
    for key, val in (strings or {}).items():
        raw = _vi(len(val)) + val
        entries[key.encode()] = _entry(DT_STRING, (), len(data), len(raw), 0)
        data += raw
    with open(f"{prefix}.data-00000-of-00001", "wb") as f:
        f.write(data)

    out = bytearray()

    def emit(block_bytes):
        off = len(out)
        out.extend(block_bytes + b"\x00")
        out.extend(struct.pack("<I", crc32c_mask(crc32c(block_bytes + b"\x00"))))
        return _vi(off) + _vi(len(block_bytes))

    index = _Block(restart_interval=1)
    blk = _Block()
    for key in sorted(entries):
        blk.add(key, entries[key])
        if len(blk.buf) >= block_size:
            index.add(key, emit(blk.finish()))
            blk = _Block()
    if blk.count:
        index.add(blk.last, emit(blk.finish()))
    meta = emit(_Block().finish())
    idx = emit(index.finish())
    footer = meta + idx
    footer += b"\x00" * (40 - len(footer))
    footer += struct.pack("<II", 0x8b80fb57, 0xdb477524)
    out.extend(footer)
    with open(f"{prefix}.index", "wb") as f:
        f.write(out)


def _crc(raw):
    from nafp_b200.model.tf_checkpoint import _crc32c_fast
    return _crc32c_fast(np.frombuffer(raw, np.uint8))


def reference_variable_names(weights):
    """The object-graph keys ``tf.train.Checkpoint(model=FingerPrinter())`` gives the variables of
    ``model/fp/nnfp.py`` -> arrays taken from a weights dict of ``model/weights.py``."""
    suf = "/.ATTRIBUTES/VARIABLE_VALUE"
    out = {}
    for i in range(8):
        base = f"model/front_conv/layer_with_weights-{i}/"
        for half, tag in (("a", "1x3"), ("b", "3x1")):
            out[f"{base}conv2d_{tag}/kernel{suf}"] = weights[f"conv{i}_{half}_w"]
            out[f"{base}conv2d_{tag}/bias{suf}"] = weights[f"conv{i}_{half}_b"]
            out[f"{base}BN_{tag}/gamma{suf}"] = weights[f"ln{i}_{half}_g"]
            out[f"{base}BN_{tag}/beta{suf}"] = weights[f"ln{i}_{half}_b"]
    for q in range(128):
        base = f"model/div_enc/split_fc_layers/{q}/layer_with_weights-"
        out[f"{base}0/kernel{suf}"] = weights["div_w1"][q]
        out[f"{base}0/bias{suf}"] = weights["div_b1"][q]
        out[f"{base}1/kernel{suf}"] = weights["div_w2"][q]
        out[f"{base}1/bias{suf}"] = weights["div_b2"][q]
    return out
