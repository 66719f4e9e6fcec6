"""Recover the parameter inventory (names + shapes) from the reference's TF checkpoint *index* files.

Run in the build container (reads /root/reference, which does not travel to the GPU box):

    python tests/golden/make_ckpt_inventory.py

writes tests/golden/ckpt_inventory.json.  This is the one structural known-answer the reference offers
for the path: the data blobs are missing, the .index SSTables survive.

Format notes (TensorFlow tensor_bundle): the .index file is a LevelDB-style table: data blocks of
prefix-compressed (key, value) entries + restart array, an index block, and a 48-byte footer holding
the metaindex and index BlockHandles (varint64 offset,size) and the magic 0xdb4775248b80fb57.  Each
value is a BundleEntryProto {1: dtype, 2: TensorShapeProto{2: Dim{1: size}}, 3: shard_id, 4: offset,
5: size, 6: crc32c}.  The entry with the empty key is the BundleHeaderProto.
"""
from __future__ import annotations

import json
import os
import sys

REF = "/root/reference/DeepSC-GAN/checkpoint"
FILES = {
    "Transeiver_Star": os.path.join(REF, "ckpt-9.index"),
    "Transeiver_star": os.path.join(REF, "FFN", "epoch-20", "ckpt-9.index"),
}


def varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def read_block(buf, offset, size):
    """Yield (key, value) of one table block (no compression in TF bundles: type byte 0)."""
    data = buf[offset:offset + size]
    assert buf[offset + size] == 0, "compressed block not supported"
    n_restarts = int.from_bytes(data[-4:], "little")
    end = len(data) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = varint(data, pos)
        non_shared, pos = varint(data, pos)
        vlen, pos = varint(data, pos)
        key = key[:shared] + data[pos:pos + non_shared]
        pos += non_shared
        yield key, data[pos:pos + vlen]
        pos += vlen


def parse_proto(buf):
    """Minimal protobuf walk -> {field: [values]} (varint and length-delimited only, fixed32 skipped)."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = varint(buf, pos)
        elif wt == 2:
            ln, pos = varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError(f"wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def entries(path):
    buf = open(path, "rb").read()
    footer = buf[-48:]
    assert int.from_bytes(footer[-8:], "little") == 0xDB4775248B80FB57, "bad table magic"
    pos = 0
    _, pos = varint(footer, pos)       # metaindex offset
    _, pos = varint(footer, pos)       # metaindex size
    idx_off, pos = varint(footer, pos)
    idx_size, pos = varint(footer, pos)
    for _, handle in read_block(buf, idx_off, idx_size):
        off, p = varint(handle, 0)
        size, p = varint(handle, p)
        yield from read_block(buf, off, size)


def inventory(path):
    inv = {}
    for key, value in entries(path):
        name = key.decode("utf-8")
        if not name or "/.ATTRIBUTES/" not in name:
            continue                                     # header, object graph, save_counter bookkeeping
        msg = parse_proto(value)
        shape = []
        if 2 in msg:
            for dim in parse_proto(msg[2][0]).get(2, []):
                shape.append(parse_proto(dim).get(1, [0])[0])
        name = name.replace("/.ATTRIBUTES/VARIABLE_VALUE", "")
        if "/.OPTIMIZER_SLOT" in name or name.endswith("save_counter") or name.startswith("optimizer"):
            continue
        inv[name] = shape
    return inv


def main():
    out = {}
    for cls, path in FILES.items():
        inv = inventory(path)
        total = 0
        for shp in inv.values():
            n = 1
            for d in shp:
                n *= d
            total += n
        out[cls] = {"source": path.replace("/root/reference/", ""), "variables": inv, "total_parameters": total}
        print(cls, len(inv), "variables", total, "parameters", file=sys.stderr)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ckpt_inventory.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", dst)


if __name__ == "__main__":
    main()
