"""On-disk interchange with a stock texrecon (INTEGRATION.md route B).

  OUT_data_costs.spt  SparseTable<uint32,uint16,float>::save_to_file / load_from_file
                      (libs/tex/sparse_table.h:112-187): ASCII header "SPT 0.2 <cols> <rows> <nnz>\\n"
                      followed by nnz packed records (u32 col = face, u16 row = view, f32 value),
                      column major, i.e. exactly the CSR-by-face order of DataCosts.
                      Accepted by `texrecon -D` (apps/texrecon/texrecon.cpp:107-117).
  OUT_labeling.vec    vector_to_file<std::size_t> (libs/tex/util.h:104-131): raw size_t[F], no header.
                      Accepted by `texrecon -L` (texrecon.cpp:137-158).
Errors mirror the reference's util::FileException messages.
"""
from __future__ import annotations

import numpy as np

_REC = np.dtype([("col", "<u4"), ("row", "<u2"), ("val", "<f4")])  # packed: 10 bytes
assert _REC.itemsize == 10


class FileException(RuntimeError):
    pass


def save_data_costs(filename, face_ptr, view, cost, num_views):
    face_ptr = np.asarray(face_ptr, np.uint64)
    F, nnz = len(face_ptr) - 1, int(face_ptr[-1])
    rec = np.empty(nnz, _REC)
    rec["col"] = np.repeat(np.arange(F, dtype=np.uint32), np.diff(face_ptr.astype(np.int64)))
    rec["row"] = np.asarray(view, np.uint16)[:nnz]
    rec["val"] = np.asarray(cost, np.float32)[:nnz]
    with open(filename, "wb") as f:
        f.write(f"SPT 0.2 {F} {int(num_views)} {nnz}\n".encode("ascii"))
        rec.tofile(f)


def load_data_costs(filename, num_faces=None, num_views=None):
    """Returns (face_ptr u64[F+1], view u16[nnz], cost f32[nnz], num_views).  Like the reference,
    a table with different dimensions than expected is rejected (sparse_table.h:166-169)."""
    with open(filename, "rb") as f:
        header = f.readline().decode("ascii", errors="replace").split()
        if len(header) < 5 or header[0] != "SPT":
            raise FileException(f"{filename}: Not a SparseTable file!")
        if header[1] != "0.2":
            raise FileException(f"{filename}: Incompatible version of SparseTable file!")
        cols, rows, nnz = int(header[2]), int(header[3]), int(header[4])
        if (num_faces is not None and cols != num_faces) or (num_views is not None and rows != num_views):
            raise FileException(f"{filename}: SparseTable has different dimension!")
        rec = np.fromfile(f, _REC, count=nnz)
    if len(rec) != nnz:
        raise FileException(f"{filename}: truncated SparseTable file")
    col = rec["col"].astype(np.int64)
    if nnz and (np.any(np.diff(col) < 0) or col.max() >= cols):
        # set_value() order is free in the reference; DataCosts written by texrecon are column major
        order = np.argsort(col, kind="stable")
        rec, col = rec[order], col[order]
    face_ptr = np.zeros(cols + 1, np.uint64)
    np.add.at(face_ptr, col + 1, 1)
    face_ptr = np.cumsum(face_ptr).astype(np.uint64)
    return face_ptr, rec["row"].copy(), rec["val"].copy(), rows


def save_labeling(filename, labels):
    np.asarray(labels).astype("<u8").tofile(filename)  # std::size_t on LP64


def load_labeling(filename, num_faces=None, num_views=None):
    """texrecon.cpp:141-153: wrong size or label > number of views aborts."""
    lab = np.fromfile(filename, "<u8")
    if num_faces is not None and len(lab) != num_faces:
        raise FileException("Wrong labeling file for this mesh/scene combination... aborting!")
    if num_views is not None and len(lab) and lab.max() > num_views:
        raise FileException("Wrong labeling file for this mesh/scene combination... aborting!")
    return lab.astype(np.uint32)
