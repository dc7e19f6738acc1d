"""TensorFlow-free reader / writer of the TensorBundle files a Keras SavedModel keeps its variables in
(``<dir>/variables/variables.index`` + ``variables.data-00000-of-00001``), so checkpoints written by the
reference's ``vae.encoder.save(dir)`` / ``vae.decoder.save(dir)`` (train.py:127-128) load into this runtime and
the ones written here can be read back with ``tf.train.load_checkpoint``.

Formats restated from TensorFlow's published sources (TF is not vendored with the reference and not installable
here, so this is **unpinned against files written by TF itself** - the tests round-trip reader and writer):
* index = a LevelDB-style sorted table: data blocks of prefix-compressed (key, value) entries with restart
  points, a metaindex block, an index block of block handles and a 48-byte footer (magic 0xdb4775248b80fb57);
  every block is followed by a 1-byte compression tag and a masked CRC32C.
* key "" -> BundleHeaderProto, every other key -> BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}.
* Keras object-graph keys of a Sequential: ``layer_with_weights-<i>/{kernel,bias}/.ATTRIBUTES/VARIABLE_VALUE``.
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, List, Tuple

import numpy as np

MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 4: np.uint8, 10: np.bool_}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}
KERAS_KEY = re.compile(r"^layer_with_weights-(\d+)/(kernel|bias)/\.ATTRIBUTES/VARIABLE_VALUE$")

# ---- CRC32C (Castagnoli), table driven (about 1 s per 10 MB in the interpreter: checkpoints are written rarely)
_POLY = 0x82F63B78
_T = np.zeros(256, np.uint32)
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _T[_i] = _c
_TL = [int(v) for v in _T]


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    t = _TL
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / minimal protobuf
def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _parse_proto(buf: bytes) -> List[Tuple[int, int, object]]:
    """[(field, wire type, value)] of one message; nested messages stay bytes."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        elif wt == 5:
            v, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.append((f, wt, v))
    return out


def _entry_proto(dtype_id: int, shape, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    msg = b"\x08" + _put_varint(dtype_id) + b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        msg += b"\x20" + _put_varint(offset)
    msg += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return msg


def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": [], "shard": 0, "offset": 0, "size": 0, "crc": None}
    for f, _, v in _parse_proto(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, dim in _parse_proto(v):
                if f2 == 2:
                    size = 0
                    for f3, _, s in _parse_proto(dim):
                        if f3 == 1:
                            size = s
                    e["shape"].append(size)
        elif f == 3:
            e["shard"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc"] = v
    return e


# ---- sorted table
def _parse_block(raw: bytes) -> List[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", raw, len(raw) - 4)[0]
    end = len(raw) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(raw, pos)
        non_shared, pos = _get_varint(raw, pos)
        vlen, pos = _get_varint(raw, pos)
        key = key[:shared] + raw[pos:pos + non_shared]
        pos += non_shared
        out.append((key, raw[pos:pos + vlen]))
        pos += vlen
    return out


def _read_block(f: bytes, offset: int, size: int, verify: bool) -> List[Tuple[bytes, bytes]]:
    raw, tag = f[offset:offset + size], f[offset + size]
    if verify:
        want = struct.unpack_from("<I", f, offset + size + 1)[0]
        if mask_crc(crc32c(f[offset:offset + size + 1])) != want:
            raise ValueError("TensorBundle index: block checksum mismatch")
    if tag != 0:
        raise NotImplementedError("TensorBundle index block is compressed (snappy); TensorFlow writes bundles uncompressed")
    return _parse_block(raw)


def read_index(path: str, verify: bool = True) -> Dict[str, dict]:
    f = open(path, "rb").read()
    if len(f) < 48 or struct.unpack_from("<Q", f, len(f) - 8)[0] != MAGIC:
        raise ValueError(f"{path}: not a TensorBundle index (bad magic)")
    foot = f[-48:]
    _, p = _get_varint(foot, 0)
    _, p = _get_varint(foot, p)
    ioff, p = _get_varint(foot, p)
    isize, p = _get_varint(foot, p)
    entries = {}
    for _, handle in _read_block(f, ioff, isize, verify):
        boff, q = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, q)
        for k, v in _read_block(f, boff, bsize, verify):
            if k == b"":
                continue                      # BundleHeaderProto
            entries[k.decode()] = _parse_entry(v)
    return entries


def read_bundle(prefix: str, verify: bool = True) -> Dict[str, np.ndarray]:
    """{checkpoint key: array} of ``<prefix>.index`` + ``<prefix>.data-*``."""
    entries = read_index(prefix + ".index", verify)
    # shard files are named exactly <prefix>.data-%05d-of-%05d; the shard count is the number of files matching that
    # pattern with a common total (stray .tempstate / editor files next to them do not shift the shard index)
    import re
    pat = re.compile(re.escape(os.path.basename(prefix)) + r"\.data-(\d{5})-of-(\d{5})$")
    byidx = {}
    for fn in os.listdir(os.path.dirname(prefix) or "."):
        m = pat.match(fn)
        if m:
            byidx.setdefault(int(m.group(2)), {})[int(m.group(1))] = fn
    total = max((t for t, d in byidx.items() if len(d) == t), default=None)
    if total is None:
        raise ValueError(f"{prefix}: no complete set of .data-XXXXX-of-XXXXX shard files")
    shards = [byidx[total][i] for i in range(total)]
    out = {}
    handles = {}
    for k, e in entries.items():
        if e["dtype"] not in _DTYPES:
            continue                          # strings (object graph proto), variants: not model weights
        name = shards[e["shard"]]
        if name not in handles:
            handles[name] = open(os.path.join(os.path.dirname(prefix), name), "rb")
        fh = handles[name]
        fh.seek(e["offset"])
        raw = fh.read(e["size"])
        if verify and e["crc"] is not None and mask_crc(crc32c(raw)) != e["crc"]:
            raise ValueError(f"{k}: tensor checksum mismatch")
        out[k] = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    for fh in handles.values():
        fh.close()
    return out


def _build_block(items: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write ``<prefix>.index`` + ``<prefix>.data-00000-of-00001`` (one shard, little endian, uncompressed)."""
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    items = [(b"", b"\x08\x01\x1a\x02\x08\x01")]          # BundleHeaderProto{num_shards: 1, version{producer: 1}}
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as data:
        for k in sorted(tensors):
            a = np.asarray(tensors[k], order="C")
            raw = a.tobytes()
            items.append((k.encode(), _entry_proto(_DTYPE_IDS[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            data.write(raw)
            offset += len(raw)
    f = bytearray()

    def emit(block: bytes) -> bytes:
        off = len(f)
        f.extend(block + b"\x00")
        f.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    data_handle = emit(_build_block(items))
    meta_handle = emit(_build_block([]))
    index_handle = emit(_build_block([(items[-1][0] + b"\x00", data_handle)], restart_interval=1))
    foot = meta_handle + index_handle
    f.extend(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", MAGIC))
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(f))


# ---- Keras Sequential <-> bundle
def keras_weights_from_bundle(prefix: str) -> List[np.ndarray]:
    """Variables of a saved Keras Sequential in ``trainable_weights`` order (layer index, kernel before bias)."""
    found = []
    for k, a in read_bundle(prefix).items():
        m = KERAS_KEY.match(k)
        if m:
            found.append((int(m.group(1)), 0 if m.group(2) == "kernel" else 1, a))
    if not found:
        raise ValueError(f"{prefix}: no layer_with_weights-*/kernel|bias variables found")
    return [a for _, _, a in sorted(found, key=lambda t: (t[0], t[1]))]


def keras_weights_to_bundle(prefix: str, weights: List[np.ndarray]) -> None:
    """Inverse of :func:`keras_weights_from_bundle` for a stack of (kernel, bias) layers."""
    assert len(weights) % 2 == 0
    t = {}
    for i in range(0, len(weights), 2):
        t[f"layer_with_weights-{i // 2}/kernel/.ATTRIBUTES/VARIABLE_VALUE"] = np.asarray(weights[i], np.float32)
        t[f"layer_with_weights-{i // 2}/bias/.ATTRIBUTES/VARIABLE_VALUE"] = np.asarray(weights[i + 1], np.float32)
    write_bundle(prefix, t)
