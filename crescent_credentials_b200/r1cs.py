"""iden3 .r1cs (v1, BN254) -> ConstraintMatrices in CSR form, built ONCE per circuit.

File format: forks/circom-compat/src/circom/r1cs_reader.rs:54-256 (worked example at :266-344).  Variable order:
CircomCircuit::generate_constraints with wire_mapping = None (forks/circom-compat/src/circom/circuit.rs:28-87,
builder.rs:64): instance i <-> wire i for i < 1 + n_pub_out + n_pub_in, witness j <-> wire num_inputs + j, so the CSR
column is the circom wire index.  As ark-relations' to_matrices() does after LC inlining, duplicate wires inside one
linear combination are summed and zero coefficients dropped.  Coefficients stay canonical little-endian words
(ENC_CANONICAL): the library converts them to Montgomery form on the GPU."""
from __future__ import annotations

import struct

import numpy as np

from . import ffi
from .groth16 import R_MOD, ConstraintMatrices

R1CS_PRIME_BYTES = bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")


class R1CSFile:
    """Header + constraints exactly as R1CSFile::new returns them (r1cs_reader.rs:47-148)."""

    def __init__(self, data: bytes):
        if data[:4] != b"r1cs":
            raise ValueError("Invalid magic number")
        self.version, nsec = struct.unpack_from("<II", data, 4)
        if self.version != 1:
            raise ValueError("Unsupported version")
        off = 12
        secs = {}
        for _ in range(nsec):
            ty, sz = struct.unpack_from("<IQ", data, off)
            off += 12
            secs[ty] = (off, sz)
            off += sz
        for ty, nm in ((1, "header"), (2, "constraint"), (3, "wire2label")):
            if ty not in secs:
                raise ValueError(f"No section offset for {nm} type found")
        ho, hs = secs[1]
        (self.field_size,) = struct.unpack_from("<I", data, ho)
        if self.field_size != 32:
            raise ValueError("This parser only supports 32-byte fields")
        if hs != 32 + self.field_size:
            raise ValueError("Invalid header section size")
        self.prime_size = bytes(data[ho + 4:ho + 36])
        if self.prime_size != R1CS_PRIME_BYTES:
            raise ValueError("This parser only supports bn256")
        (self.n_wires, self.n_pub_out, self.n_pub_in, self.n_prv_in, self.n_labels,
         self.n_constraints) = struct.unpack_from("<IIIIQI", data, ho + 36)
        mo, ms = secs[3]
        if ms != self.n_wires * 8:
            raise ValueError("Invalid map section size")
        self.wire_mapping = np.frombuffer(data, dtype="<u8", count=self.n_wires, offset=mo)
        if self.n_wires and self.wire_mapping[0] != 0:
            raise ValueError("Wire 0 should always be mapped to 0")
        self._data = data
        self._cons_off = secs[2][0]

    def constraints(self):
        """Yields (A, B, C), each a list of (wire, coeff_int) -- r1cs_reader.rs:209-236."""
        data, p = self._data, self._cons_off
        for _ in range(self.n_constraints):
            row = []
            for _k in range(3):
                (nv,) = struct.unpack_from("<I", data, p)
                p += 4
                vec = []
                for _j in range(nv):
                    (w,) = struct.unpack_from("<I", data, p)
                    vec.append((w, int.from_bytes(data[p + 4:p + 36], "little")))
                    p += 36
                row.append(vec)
            yield tuple(row)


def _load_matrices_rowwise(data: bytes) -> ConstraintMatrices:
    """Term-by-term restatement (the definition the vectorised loader below is tested against)."""
    f = R1CSFile(data)
    num_inputs = 1 + f.n_pub_in + f.n_pub_out  # r1cs_reader.rs:27
    rp = [np.zeros(f.n_constraints + 1, dtype=np.uint64) for _ in range(3)]
    cols = [[], [], []]
    vals = [bytearray(), bytearray(), bytearray()]
    for i, con in enumerate(f.constraints()):
        for k in range(3):
            acc = {}
            for w, v in con[k]:
                if w >= f.n_wires:
                    raise ValueError("wire index out of range")
                acc[w] = (acc.get(w, 0) + v) % R_MOD
            for w in sorted(acc):
                if acc[w]:
                    cols[k].append(w)
                    vals[k] += acc[w].to_bytes(32, "little")
            rp[k][i + 1] = len(cols[k])
    col = [np.array(c, dtype=np.uint32) for c in cols]
    val = [np.frombuffer(bytes(v), dtype="<u8").reshape(-1, 4).copy() for v in vals]
    return ConstraintMatrices(num_inputs, f.n_wires - num_inputs, f.n_constraints, rp, col, val, ffi.ENC_CANONICAL)


def load_matrices(data: bytes) -> ConstraintMatrices:
    """.r1cs -> CSR.  The only sequential part of the format is finding where each linear combination starts (every count
    sits behind the previous terms): one pass over the 3 * n_constraints counts; everything else is array work -- terms are
    gathered with one fancy index per matrix, rows are sorted by wire with one lexsort, and only rows that repeat a wire or
    hold a zero / unreduced coefficient (rare in circom output) are merged term by term."""
    f = R1CSFile(data)
    num_inputs = 1 + f.n_pub_in + f.n_pub_out  # r1cs_reader.rs:27
    nc = f.n_constraints
    buf = np.frombuffer(data, dtype=np.uint8)
    # pass 1: counts and byte offsets of the 3 * nc linear combinations
    counts = np.zeros(3 * nc, dtype=np.int64)
    starts = np.zeros(3 * nc, dtype=np.int64)
    p = f._cons_off
    unpack = struct.Struct("<I").unpack_from
    for t in range(3 * nc):
        (nv,) = unpack(data, p)
        counts[t] = nv
        starts[t] = p + 4
        p += 4 + 36 * nv
    if p > len(data):
        raise ValueError("constraint section runs past the end of the file")
    rp, col, val = [], [], []
    for k in range(3):
        cnt, st = counts[k::3], starts[k::3]
        ptr = np.zeros(nc + 1, dtype=np.int64)
        np.cumsum(cnt, out=ptr[1:])
        nnz = int(ptr[-1])
        row = np.repeat(np.arange(nc, dtype=np.int64), cnt)
        # byte offset of every term: start of its linear combination + 36 * (index inside it)
        off = np.repeat(st, cnt) + 36 * (np.arange(nnz, dtype=np.int64) - np.repeat(ptr[:-1], cnt))
        rec = buf[(off[:, None] + np.arange(36, dtype=np.int64)[None, :]).reshape(-1)].reshape(nnz, 36) if nnz else np.zeros((0, 36), np.uint8)
        wires = rec[:, :4].copy().view("<u4").reshape(-1).astype(np.uint32)
        coeff = rec[:, 4:].copy().view("<u8").reshape(-1, 4)
        if nnz and wires.max() >= f.n_wires:
            raise ValueError("wire index out of range")
        order = np.lexsort((wires, row))  # stable: by row, then by wire
        row, wires, coeff = row[order], wires[order], coeff[order]
        # rows that need term-by-term treatment: a repeated wire, a zero coefficient, or a coefficient >= r
        dup = np.zeros(nnz, dtype=bool)
        if nnz > 1:
            dup[1:] = (row[1:] == row[:-1]) & (wires[1:] == wires[:-1])
        zero = ~coeff.any(axis=1) if nnz else np.zeros(0, dtype=bool)
        big = np.zeros(nnz, dtype=bool)
        if nnz:
            rl = np.array([(R_MOD >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
            ge = np.zeros(nnz, dtype=bool)
            eq = np.ones(nnz, dtype=bool)
            for i in (3, 2, 1, 0):
                ge |= eq & (coeff[:, i] > rl[i])
                eq &= coeff[:, i] == rl[i]
            big = ge | eq
        bad_rows = np.unique(row[dup | zero | big])
        if bad_rows.size:
            keep = ~np.isin(row, bad_rows)
            extra_row, extra_wire, extra_val = [], [], []
            for i in bad_rows:
                sel = np.nonzero(row == i)[0]
                acc = {}
                for j in sel:
                    w = int(wires[j])
                    acc[w] = (acc.get(w, 0) + sum(int(coeff[j, q]) << (64 * q) for q in range(4))) % R_MOD
                for w in sorted(acc):
                    if acc[w]:
                        extra_row.append(int(i))
                        extra_wire.append(w)
                        extra_val.append([(acc[w] >> (64 * q)) & 0xFFFFFFFFFFFFFFFF for q in range(4)])
            row = np.concatenate([row[keep], np.array(extra_row, dtype=np.int64)])
            wires = np.concatenate([wires[keep], np.array(extra_wire, dtype=np.uint32)])
            coeff = np.concatenate([coeff[keep], np.array(extra_val, dtype=np.uint64).reshape(-1, 4)])
            order = np.lexsort((wires, row))
            row, wires, coeff = row[order], wires[order], coeff[order]
        out_ptr = np.zeros(nc + 1, dtype=np.uint64)
        np.cumsum(np.bincount(row, minlength=nc), out=out_ptr[1:])
        rp.append(out_ptr)
        col.append(np.ascontiguousarray(wires, dtype=np.uint32))
        val.append(np.ascontiguousarray(coeff, dtype=np.uint64))
    return ConstraintMatrices(num_inputs, f.n_wires - num_inputs, nc, rp, col, val, ffi.ENC_CANONICAL)
