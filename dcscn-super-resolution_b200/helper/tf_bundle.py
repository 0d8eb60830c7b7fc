"""
TensorFlow V2 "tensor bundle" checkpoint reader / writer (no TensorFlow needed).

Replaces what `tf.train.Saver.restore / save` do for the reference
(helper/tf_graph.py:263-296, `load_model` / `save_model`).  The on-disk format
is the one the reference's shipped `models/*.ckpt.{index,data-00000-of-00001}`
files use:

  <prefix>.index                 LevelDB-style SSTable, uncompressed blocks.
      key ""            -> BundleHeaderProto {1:num_shards, 2:endianness, 3:VersionDef{1:producer}}
      key <var name>    -> BundleEntryProto  {1:dtype, 2:TensorShapeProto{2:dim{1:size}},
                                              3:shard_id, 4:offset, 5:size, 6:crc32c (fixed32, masked)}
  <prefix>.data-00000-of-00001   raw little-endian tensor bytes at [offset, offset+size)

Only float32 tensors are produced by the reference graph (dtype enum 1).
"""

import os
import struct

import numpy as np

_TABLE_MAGIC = 0xdb4775248b80fb57
_DT_FLOAT = 1
_BLOCK_RESTART_INTERVAL = 16
_MASK_DELTA = 0xa282ead8

# ---------------------------------------------------------------- crc32c ----

_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        table = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if (c & 1) else (c >> 1)
            table[i] = c
        # slicing-by-8 tables for a vectorised-enough pure python loop
        tables = [table]
        for k in range(1, 8):
            prev = tables[-1]
            tables.append((prev >> 8) ^ table[prev & 0xFF])
        _CRC_TABLE = [t.tolist() for t in tables]
    return _CRC_TABLE


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli) of `data` (bytes-like)."""
    t = _crc_table()
    t0, t1, t2, t3, t4, t5, t6, t7 = t
    crc ^= 0xFFFFFFFF
    mv = memoryview(data).cast("B")
    n = len(mv)
    n8 = n - (n % 8)
    if n8:
        words = np.frombuffer(mv[:n8], dtype="<u4").tolist()
        for i in range(0, len(words), 2):
            lo = words[i] ^ crc
            hi = words[i + 1]
            crc = (t7[lo & 0xFF] ^ t6[(lo >> 8) & 0xFF] ^ t5[(lo >> 16) & 0xFF] ^ t4[lo >> 24] ^
                   t3[hi & 0xFF] ^ t2[(hi >> 8) & 0xFF] ^ t1[(hi >> 16) & 0xFF] ^ t0[hi >> 24])
    for b in mv[n8:]:
        crc = t0[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    """TF/LevelDB 'masked' crc: rotate right by 15 and add a constant."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# --------------------------------------------------------------- varints ----

def _get_varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


# ------------------------------------------------------------ mini-proto ----

def _parse_proto(buf):
    """Returns a list of (field_number, wire_type, value)."""
    pos = 0
    fields = []
    n = len(buf)
    while pos < n:
        tag, pos = _get_varint(buf, pos)
        fn, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        fields.append((fn, wt, v))
    return fields


def _parse_entry(buf):
    entry = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
    for fn, _, v in _parse_proto(buf):
        if fn == 1:
            entry["dtype"] = v
        elif fn == 2:
            dims = []
            for fn2, _, v2 in _parse_proto(v):
                if fn2 == 2:
                    size = 0
                    for fn3, _, v3 in _parse_proto(v2):
                        if fn3 == 1:
                            size = v3
                    dims.append(size)
            entry["shape"] = dims
        elif fn == 3:
            entry["shard_id"] = v
        elif fn == 4:
            entry["offset"] = v
        elif fn == 5:
            entry["size"] = v
        elif fn == 6:
            entry["crc32c"] = v
    return entry


def _encode_entry(shape, offset, size, crc):
    shape_proto = b""
    for d in shape:
        dim = b"\x08" + _put_varint(int(d))
        shape_proto += b"\x12" + _put_varint(len(dim)) + dim
    out = b"\x08" + _put_varint(_DT_FLOAT)
    out += b"\x12" + _put_varint(len(shape_proto)) + shape_proto
    # shard_id 0 is the proto default and is omitted, like TF does
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size)
    out += b"\x35" + struct.pack("<I", crc)
    return out


# --------------------------------------------------------------- sstable ----

def _read_block(data, offset, size):
    """Decode one uncompressed SSTable block into [(key, value)]."""
    block = data[offset:offset + size]
    block_type = data[offset + size]
    if block_type != 0:
        raise ValueError("compressed SSTable blocks are not supported (type %d)" % block_type)
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos = 0
    key = b""
    out = []
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        value_len, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        value = bytes(block[pos:pos + value_len])
        pos += value_len
        out.append((key, value))
    return out


class BundleReader:
    """Reads `<prefix>.index` + `<prefix>.data-00000-of-00001`."""

    def __init__(self, prefix):
        self.prefix = prefix
        index_path = prefix + ".index"
        if not os.path.isfile(index_path):
            raise FileNotFoundError(index_path)
        with open(index_path, "rb") as f:
            data = f.read()
        if len(data) < 48:
            raise ValueError("index file too small")
        magic = struct.unpack_from("<Q", data, len(data) - 8)[0]
        if magic != _TABLE_MAGIC:
            raise ValueError("bad SSTable magic in %s" % index_path)
        footer = data[-48:]
        pos = 0
        _, pos = _get_varint(footer, pos)  # metaindex offset
        _, pos = _get_varint(footer, pos)  # metaindex size
        index_off, pos = _get_varint(footer, pos)
        index_size, pos = _get_varint(footer, pos)

        self.header = None
        self.entries = {}
        for _, handle in _read_block(data, index_off, index_size):
            off, p = _get_varint(handle, 0)
            size, p = _get_varint(handle, p)
            for key, value in _read_block(data, off, size):
                if key == b"":
                    self.header = _parse_proto(value)
                else:
                    self.entries[key.decode("utf-8")] = _parse_entry(value)
        self.num_shards = 1
        if self.header:
            for fn, _, v in self.header:
                if fn == 1:
                    self.num_shards = v
        self._data = None

    def keys(self):
        return sorted(self.entries.keys())

    def has_tensor(self, name):
        return name in self.entries

    def shape(self, name):
        return list(self.entries[name]["shape"])

    def _data_file(self, shard_id):
        if self._data is None:
            self._data = {}
        if shard_id not in self._data:
            path = "%s.data-%05d-of-%05d" % (self.prefix, shard_id, self.num_shards)
            self._data[shard_id] = np.memmap(path, dtype=np.uint8, mode="r")
        return self._data[shard_id]

    def get_tensor(self, name, verify_crc=False):
        e = self.entries[name]
        if e["dtype"] != _DT_FLOAT:
            raise ValueError("tensor %s: only float32 is supported (dtype %d)" % (name, e["dtype"]))
        raw = self._data_file(e["shard_id"])[e["offset"]:e["offset"] + e["size"]]
        if verify_crc and e["crc32c"] is not None:
            if masked_crc32c(raw.tobytes()) != e["crc32c"]:
                raise ValueError("crc32c mismatch for tensor %s" % name)
        arr = np.frombuffer(raw.tobytes(), dtype="<f4").astype(np.float32)
        return arr.reshape(e["shape"]) if e["shape"] else arr.reshape(())


def _build_block(items):
    """Encode sorted [(key, value)] into one SSTable block (without trailer)."""
    out = bytearray()
    restarts = []
    last_key = b""
    for i, (key, value) in enumerate(items):
        if i % _BLOCK_RESTART_INTERVAL == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            m = min(len(last_key), len(key))
            while shared < m and last_key[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        out += key[shared:] + value
        last_key = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _block_trailer(block):
    crc = crc32c(block + b"\x00")
    masked = (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF
    return b"\x00" + struct.pack("<I", masked)


def write_bundle(prefix, tensors, block_size=4096):
    """
    Writes {name: float32 ndarray} as a TF V2 bundle at `prefix` (the
    reference's `saver.save(sess, filename)`, helper/tf_graph.py:291).
    Keys are stored in bytewise-sorted order exactly like TF's BundleWriter.
    """
    directory = os.path.dirname(prefix)
    if directory:
        os.makedirs(directory, exist_ok=True)

    names = sorted(tensors.keys(), key=lambda s: s.encode("utf-8"))
    entries = []
    offset = 0
    # both files are written under temporary names and renamed into place, the index LAST: a reader (or a crash)
    # never sees an index that points into a half-written data file
    tmp_tag = ".tmp%d" % os.getpid()
    with open(prefix + ".data-00000-of-00001" + tmp_tag, "wb") as f:
        for name in names:
            src = np.asarray(tensors[name], dtype="<f4")
            shape = src.shape                      # () for scalars such as beta1_power (ascontiguousarray would make it (1,))
            raw = np.ascontiguousarray(src).tobytes()
            f.write(raw)
            entries.append((name.encode("utf-8"),
                            _encode_entry(shape, offset, len(raw), masked_crc32c(raw))))
            offset += len(raw)

    # BundleHeaderProto: num_shards=1, endianness=LITTLE(0, omitted), version{producer=1}
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"
    items = [(b"", header)] + entries

    out = bytearray()
    index_items = []
    cur = []
    cur_size = 0

    def flush():
        nonlocal cur, cur_size
        if not cur:
            return
        block = _build_block(cur)
        handle = _put_varint(len(out)) + _put_varint(len(block))
        # index key: any key >= last key of this block and < first key of next; use last key
        index_items.append((cur[-1][0], handle))
        out.extend(block)
        out.extend(_block_trailer(block))
        cur = []
        cur_size = 0

    for key, value in items:
        cur.append((key, value))
        cur_size += len(key) + len(value) + 3
        if cur_size >= block_size:
            flush()
    flush()

    meta_block = _build_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta_block))
    out.extend(meta_block)
    out.extend(_block_trailer(meta_block))

    index_block = _build_block(index_items)
    index_handle = _put_varint(len(out)) + _put_varint(len(index_block))
    out.extend(index_block)
    out.extend(_block_trailer(index_block))

    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer))
    footer += struct.pack("<Q", _TABLE_MAGIC)
    out.extend(footer)

    with open(prefix + ".index" + tmp_tag, "wb") as f:
        f.write(bytes(out))
    os.replace(prefix + ".data-00000-of-00001" + tmp_tag, prefix + ".data-00000-of-00001")
    os.replace(prefix + ".index" + tmp_tag, prefix + ".index")
