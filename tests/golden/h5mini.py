"""Test infrastructure: a tiny writer of netCDF-4-shaped HDF5 files in the OLD on-disk style
(superblock 0, version-1 object headers, symbol-table groups, contiguous data) -- what h5py /
h5netcdf and old netCDF libraries produce.  The reader in ``smmregrid_b200/nc4.py`` is pinned on
files written by the real netCDF library (new style: superblock 2, version-2 headers, fractal
heaps; ``tests/golden/data/*.nc``); this writer lets the tests produce WEIGHT files in the
container (no netCDF4 / h5py here) and covers the reader's old-style branches.

    write(path, dims={"num_links": 10, ...}, variables={"src_address": (("num_links",), array), ...},
          coords={"lev": array}, attrs={"map_method": "Conservative remapping"})
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    be = 1 if dt.byteorder == ">" else 0
    if dt.kind in "iu":
        bits0 = be | (8 if dt.kind == "i" else 0)
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        n = dt.itemsize
        exp = {4: (23, 8, 0, 23, 127), 8: (52, 11, 0, 52, 1023)}[n]
        return struct.pack("<BBBBI", 0x11, be | 0x20, 8 * n - 1, 0, n) + \
            struct.pack("<HHBBBBI", 0, 8 * n, exp[0], exp[1], exp[2], exp[3], exp[4])
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0, 0, 0, dt.itemsize)
    raise TypeError(dt)


_VLEN_REF = struct.pack("<BBBBI", 0x19, 0, 0, 0, 16) + struct.pack("<BBBBI", 0x17, 0, 0, 0, 8)


def _space_msg(shape, unlimited=False) -> bytes:
    rank = len(shape)
    out = struct.pack("<BBBBI", 1, rank, 1 if unlimited else 0, 0, 0)
    out += b"".join(struct.pack("<Q", int(n)) for n in shape)
    if unlimited:
        out += b"".join(struct.pack("<Q", UNDEF) for _ in shape)
    return out


def _msg(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHBBBB", mtype, len(body), 0, 0, 0, 0) + body


def _attr_msg(name: str, dtmsg: bytes, spmsg: bytes, data: bytes) -> bytes:
    nm = name.encode() + b"\x00"
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtmsg), len(spmsg)) + _pad8(nm) + _pad8(dtmsg) + _pad8(spmsg) + data
    return _msg(0x0C, body)


def _attr(name, value) -> bytes:
    if isinstance(value, str):
        raw = value.encode() + b"\x00"
        return _attr_msg(name, _dtype_msg(np.dtype(f"S{len(raw)}")), _space_msg(()), raw)
    a = np.asarray(value)
    if a.dtype.kind == "i" and a.dtype.itemsize == 8:
        a = a.astype(np.int32)
    return _attr_msg(name, _dtype_msg(a.dtype), _space_msg(a.shape if a.ndim else ()), a.tobytes())


def _header(msgs) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\x00" * 4 + body


def write(path, dims: dict, variables: dict, coords: dict | None = None, attrs: dict | None = None,
          var_attrs: dict | None = None):
    """`variables`: name -> (dims tuple, array); `coords`: dim name -> 1-D coordinate values
    (a dimension scale that is also a variable); dims without a coordinate become bare scales."""
    coords = coords or {}
    var_attrs = var_attrs or {}
    names = list(dims) + [k for k in variables if k not in dims]
    # ---- pass 1: sizes.  Layout: superblock | root header | heap | heap data | btree | snod | gcol | objects | data
    def build(addr_of, data_addr_of, gcol_addr):
        objs = {}
        for dimid, (d, n) in enumerate(dims.items()):
            c = coords.get(d)
            arr = np.asarray(c) if c is not None else None
            dt = arr.dtype if arr is not None else np.dtype(">f4")
            msgs = [_msg(0x01, _space_msg((n,))), _msg(0x03, _dtype_msg(dt)),
                    _msg(0x05, struct.pack("<BBBB", 2, 2, 0, 0)),
                    _msg(0x08, struct.pack("<BBQQ", 3, 1, data_addr_of.get(d, UNDEF) if arr is not None else UNDEF,
                                           (arr.nbytes if arr is not None else 0))),
                    _attr("CLASS", "DIMENSION_SCALE"),
                    _attr("NAME", d if arr is not None else
                          "This is a netCDF dimension but not a netCDF variable.%10d" % n),
                    _attr("_Netcdf4Dimid", np.int32(dimid))]
            for k, v in var_attrs.get(d, {}).items():
                msgs.append(_attr(k, v))
            objs[d] = (_header(msgs), arr)
        for name, (vd, arr) in variables.items():
            arr = np.asarray(arr).copy(order="C")
            msgs = [_msg(0x01, _space_msg(arr.shape)), _msg(0x03, _dtype_msg(arr.dtype)),
                    _msg(0x05, struct.pack("<BBBB", 2, 2, 0, 0)),
                    _msg(0x08, struct.pack("<BBQQ", 3, 1, data_addr_of.get(name, UNDEF), arr.nbytes))]
            if vd:
                dl = b"".join(struct.pack("<IQI", 1, gcol_addr, 1 + list(dims).index(d)) for d in vd)
                msgs.append(_attr_msg("DIMENSION_LIST", _VLEN_REF, _space_msg((len(vd),)), dl))
            for k, v in var_attrs.get(name, {}).items():
                msgs.append(_attr(k, v))
            objs[name] = (_header(msgs), arr)
        return objs

    sorted_names = sorted(names)
    heap_data = b"\x00" * 8
    name_off = {}
    for n in sorted_names:
        name_off[n] = len(heap_data)
        heap_data += _pad8(n.encode() + b"\x00")
    root_msgs_len = 16 + 8 + 16 + sum(len(_attr(k, v)) for k, v in (attrs or {}).items())
    p = 96
    root_addr = p
    p += root_msgs_len
    heap_addr = p
    p += 32
    heap_data_addr = p
    p += len(heap_data)
    btree_addr = p
    p += 8 + 16 + 8 + 8 + 8
    snod_addr = p
    p += 8 + 40 * len(names)
    gcol_addr = p
    gcol_size = max(4096, 16 + 24 * (len(dims) + 1))
    p += gcol_size
    objs = build({}, {}, gcol_addr)
    addr_of = {}
    for n in names:
        addr_of[n] = p
        p += len(objs[n][0])
    data_addr_of = {}
    for n in names:
        arr = objs[n][1]
        if arr is not None and arr.nbytes:
            p += -p % 8
            data_addr_of[n] = p
            p += arr.nbytes
    eof = p
    objs = build(addr_of, data_addr_of, gcol_addr)

    out = bytearray(eof)
    sb = b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 64, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    out[0:len(sb)] = sb
    root = _header([_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))] + [_attr(k, v) for k, v in (attrs or {}).items()])
    assert len(root) == root_msgs_len, (len(root), root_msgs_len)
    out[root_addr:root_addr + len(root)] = root
    out[heap_addr:heap_addr + 32] = b"HEAP" + bytes(4) + struct.pack("<QQQ", len(heap_data), 1, heap_data_addr)
    out[heap_data_addr:heap_data_addr + len(heap_data)] = heap_data
    bt = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF)
    bt += struct.pack("<QQQ", 0, snod_addr, name_off[sorted_names[-1]])
    out[btree_addr:btree_addr + len(bt)] = bt
    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in sorted_names:
        sn += struct.pack("<QQII", name_off[n], addr_of[n], 0, 0) + bytes(16)
    out[snod_addr:snod_addr + len(sn)] = sn
    gc = b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", gcol_size)
    for i, d in enumerate(dims):
        gc += struct.pack("<HHIQ", i + 1, 1, 0, 8) + struct.pack("<Q", addr_of[d])
    gc += struct.pack("<HHIQ", 0, 0, 0, gcol_size - len(gc))
    out[gcol_addr:gcol_addr + len(gc)] = gc
    for n in names:
        h, arr = objs[n]
        out[addr_of[n]:addr_of[n] + len(h)] = h
        if n in data_addr_of:
            out[data_addr_of[n]:data_addr_of[n] + arr.nbytes] = arr.tobytes()
    with open(path, "wb") as f:
        f.write(bytes(out))


def write_weights(path, w, links_dim="num_links", level_dim="depth_full"):
    """A CDO-style SCRIP weights file (2-D, or the reference's padded 3-D layout with
    `link_length`, cdogenerate.py:310-343) from a CdoWeights, as netCDF-4."""
    three_d = "link_length" in w.vars
    sa = np.asarray(w["src_address"])
    rm = np.asarray(w["remap_matrix"])
    n_src, n_dst = int(w.sizes["src_grid_size"]), int(w.sizes["dst_grid_size"])
    dims = {"src_grid_size": n_src, "dst_grid_size": n_dst, links_dim: sa.shape[-1], "num_wgts": rm.shape[-1],
            "src_grid_rank": np.atleast_1d(w["src_grid_dims"]).size, "dst_grid_rank": np.atleast_1d(w["dst_grid_dims"]).size}
    lev, coords = (), {}
    if three_d:
        dims = {level_dim: int(np.asarray(w["link_length"]).size), **dims}
        lev = (level_dim,)
        coords[level_dim] = np.asarray(w.levels, dtype=np.float64)
    spec = {"src_address": lev + (links_dim,), "dst_address": lev + (links_dim,),
            "remap_matrix": lev + (links_dim, "num_wgts"), "src_grid_imask": lev + ("src_grid_size",),
            "dst_grid_imask": lev + ("dst_grid_size",), "dst_grid_frac": lev + ("dst_grid_size",),
            "src_grid_dims": ("src_grid_rank",), "dst_grid_dims": ("dst_grid_rank",),
            "dst_grid_center_lat": ("dst_grid_size",), "dst_grid_center_lon": ("dst_grid_size",)}
    variables = {}
    for name, vd in spec.items():
        a = np.asarray(w[name])
        variables[name] = (vd, a.reshape(tuple(dims[d] for d in vd)))
    if three_d:
        variables["link_length"] = (lev, np.asarray(w["link_length"]).astype(np.int32))
    write(path, dims, variables, coords=coords,
          attrs={"source_grid": str(w.attrs.get("source_grid", "src")), "dest_grid": str(w.attrs.get("dest_grid", "dst")),
                 "map_method": "Conservative remapping", "normalization": "fracarea"},
          var_attrs={"remap_matrix": {"units": "1"}})
