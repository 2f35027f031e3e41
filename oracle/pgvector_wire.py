"""CPU oracle for pgvector's wire formats -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of the bench tools may import this
module; the product (``outline_rag_b200``) never does.

PARITY UNPINNED, like oracle/cosine_topk.py: the formats belong to PostgreSQL and pgvector, both
absent from /root/reference and from this image, and the reference holds no fixtures for them.

What is restated here (published formats / algorithms):

* PostgreSQL ``COPY ... (FORMAT binary)`` file format (PostgreSQL manual, "COPY -- Binary Format"):
  11-byte signature ``PGCOPY\\n\\377\\r\\n\\0``, int32 flags (bit 16 = OIDs, bits 17-31 critical),
  int32 header-extension length + extension; per tuple an int16 field count, per field an int32
  byte length (-1 = NULL) and the type's binary ``send`` image; an int16 -1 trailer.  All big-endian.
* ``uuid_send``: the 16 bytes, most significant first.
* pgvector ``vector_send`` / ``vector_recv`` (src/vector.c [UPSTREAM]): int16 dim, int16 unused
  (must be 0), dim float4 in network byte order; vector_recv rejects dim < 1 or > 16000, a non-zero
  ``unused``, NaN and infinite elements.
* pgvector ``vector_in`` (src/vector.c [UPSTREAM]): optional whitespace, ``[``, elements parsed with
  ``strtof`` (ONE correctly rounded decimal -> fp32 conversion) separated by ``,``, ``]``, optional
  whitespace; errors for empty vectors, junk, out-of-range, NaN and infinity.

Reference anchors: the table these tuples come from is ``langchain_pg_embedding(langchain_id UUID,
embedding vector(1024))`` (reference app/database.py:118-131); the reference's driver is psycopg 3
(requirements.txt:5); the text form is what langchain-postgres sends (``str(list_of_floats)``).
"""
from __future__ import annotations

import struct
from fractions import Fraction

import numpy as np

SIGNATURE = b"PGCOPY\n\xff\r\n\x00"
VECTOR_MAX_DIM = 16000
_WS = " \t\n\r\v\f"


class WireError(ValueError):
    """kind: 'syntax' | 'dim' | 'nonfinite' | 'range' | 'format' (maps onto the ORX_ERR_* classes)."""

    def __init__(self, kind: str, msg: str):
        super().__init__(msg)
        self.kind = kind


# ----------------------------------------------------------------------------- binary
def vector_send(x: np.ndarray) -> bytes:
    x = np.ascontiguousarray(x, np.float32)
    return struct.pack(">hh", x.shape[0], 0) + x.astype(">f4").tobytes()


def vector_recv(buf: bytes, expected_dim: int | None = None) -> np.ndarray:
    if len(buf) < 4:
        raise WireError("format", "insufficient data left in message")
    dim, unused = struct.unpack(">hh", buf[:4])
    if dim < 1:
        raise WireError("dim", "vector must have at least 1 dimension")
    if dim > VECTOR_MAX_DIM:
        raise WireError("dim", f"vector cannot have more than {VECTOR_MAX_DIM} dimensions")
    if expected_dim is not None and dim != expected_dim:
        raise WireError("dim", f"expected {expected_dim} dimensions, not {dim}")
    if unused != 0:
        raise WireError("format", f"expected unused to be 0, not {unused}")
    if len(buf) != 4 + 4 * dim:
        raise WireError("format", "incorrect binary data format")
    x = np.frombuffer(buf, dtype=">f4", count=dim, offset=4).astype(np.float32)
    if np.isnan(x).any():
        raise WireError("nonfinite", "NaN not allowed in vector")
    if np.isinf(x).any():
        raise WireError("nonfinite", "infinite value not allowed in vector")
    return x


def id_to_bytes(hi: int, lo: int) -> bytes:
    return struct.pack(">QQ", int(hi), int(lo))


def copy_binary_stream(ids: np.ndarray, X, null_rows=(), header_extension: bytes = b"", trailer: bool = True,
                       flags: int = 0) -> bytes:
    """What `COPY (SELECT langchain_id, embedding ...) TO STDOUT (FORMAT binary)` emits for these rows.
    ids: uint64 [n,2] (hi, lo); X: fp32 [n, dim] (rows listed in `null_rows` have a NULL embedding)."""
    out = [SIGNATURE, struct.pack(">ii", flags, len(header_extension)), header_extension]
    nulls = set(int(i) for i in null_rows)
    for i in range(len(ids)):
        out.append(struct.pack(">hi", 2, 16))
        out.append(id_to_bytes(ids[i][0], ids[i][1]))
        if i in nulls:
            out.append(struct.pack(">i", -1))
        else:
            v = vector_send(X[i])
            out.append(struct.pack(">i", len(v)))
            out.append(v)
    if trailer:
        out.append(struct.pack(">h", -1))
    return b"".join(out)


def copy_binary_parse(stream: bytes, expected_dim: int = 1024):
    """Decode such a stream: (ids uint64 [m,2], X fp32 [m,dim], n_null).  Row order preserved."""
    if len(stream) < 19:
        raise WireError("format", "invalid COPY file header (missing length)")
    if stream[:11] != SIGNATURE:
        raise WireError("format", "COPY file signature not recognized")
    flags, ext = struct.unpack(">ii", stream[11:19])
    if flags & (1 << 16):
        raise WireError("format", "invalid COPY file header (WITH OIDS)")
    if (flags & 0xFFFFFFFF) >> 17:
        raise WireError("format", "unrecognized critical flags in COPY file header")
    if ext < 0 or len(stream) < 19 + ext:
        raise WireError("format", "invalid COPY file header (missing length)")
    p = 19 + ext
    ids, rows, n_null, done = [], [], 0, False
    while p < len(stream):
        if done:
            raise WireError("format", "received copy data after EOF marker")
        if len(stream) - p < 2:
            raise WireError("format", "unexpected EOF in COPY data")
        (nf,) = struct.unpack(">h", stream[p:p + 2])
        p += 2
        if nf == -1:
            done = True
            continue
        if nf != 2:
            raise WireError("format", f"COPY row has {nf} columns, expected 2")
        fields = []
        for _ in range(2):
            if len(stream) - p < 4:
                raise WireError("format", "unexpected EOF in COPY data")
            (ln,) = struct.unpack(">i", stream[p:p + 4])
            p += 4
            if ln == -1:
                fields.append(None)
                continue
            if ln < 0:
                raise WireError("format", "invalid field size")
            if len(stream) - p < ln:
                raise WireError("format", "unexpected EOF in COPY data")
            fields.append(stream[p:p + ln])
            p += ln
        if fields[0] is None or len(fields[0]) != 16:
            raise WireError("format", "bad uuid field")
        if fields[1] is None:
            n_null += 1
            continue
        rows.append(vector_recv(fields[1], expected_dim))
        ids.append(struct.unpack(">QQ", fields[0]))
    ida = np.array(ids, dtype=np.uint64).reshape(-1, 2)
    Xa = np.stack(rows) if rows else np.zeros((0, expected_dim), np.float32)
    return ida, Xa, n_null


# ----------------------------------------------------------------------------- text
def _f32_exact(fr: Fraction, negative: bool = False) -> np.float32:
    """Correctly rounded (nearest, ties to even) fp32 of an exact rational: what strtof returns
    (`negative` carries the sign of a "-0" token, which a rational cannot)."""
    sign = -1 if (fr < 0 or negative) else 1
    if fr == 0:
        return np.float32(-0.0) if sign < 0 else np.float32(0.0)
    a = abs(fr)
    # candidates around the double approximation; exact comparison decides
    with np.errstate(over="ignore"):
        c = np.float32(float(a)) if a < Fraction(2) ** 1000 else np.float32(np.inf)
        cands = {c, np.nextafter(c, np.float32(np.inf)), np.nextafter(c, np.float32(0.0))}
    best, best_err = None, None
    for v in cands:
        if np.isinf(v):
            # overflow threshold of round-to-nearest: values >= FLT_MAX + ulp/2 round to infinity
            vf = Fraction(2) ** 128
        else:
            vf = Fraction(float(v))
        err = abs(vf - a)
        even = np.isinf(v) or (int(np.float32(v).view(np.uint32)) & 1) == 0
        if best is None or err < best_err or (err == best_err and even and not best_even):
            best, best_err, best_even = v, err, even
    return np.float32(sign) * np.float32(best)


def vector_in(text: str, expected_dim: int | None = None) -> np.ndarray:
    """pgvector's text input.  Decimal and exponent notation (what str(float) produces), 'nan' / 'inf'
    tokens (rejected like vector_in does after strtof accepted them)."""
    import re
    s = text
    i = 0

    def skip():
        nonlocal i
        while i < len(s) and s[i] in _WS:
            i += 1

    def syntax():
        return WireError("syntax", f'invalid input syntax for type vector: "{text[:64]}"')

    skip()
    if i >= len(s) or s[i] != "[":
        raise syntax()
    i += 1
    skip()
    if i < len(s) and s[i] == "]":
        raise WireError("dim", "vector must have at least 1 dimension")
    vals = []
    num = re.compile(r"[+-]?(?:(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?|inf(?:inity)?|nan)", re.I)
    while True:
        if len(vals) == VECTOR_MAX_DIM:
            raise WireError("dim", f"vector cannot have more than {VECTOR_MAX_DIM} dimensions")
        skip()
        if i >= len(s):
            raise syntax()
        m = num.match(s, i)
        if not m:
            raise syntax()
        tok = m.group(0).lower()
        if "nan" in tok:
            raise WireError("nonfinite", "NaN not allowed in vector")
        if "inf" in tok:
            raise WireError("nonfinite", "infinite value not allowed in vector")
        v = _f32_exact(Fraction(tok), tok.startswith("-"))
        if np.isinf(v):
            raise WireError("range", f'"{m.group(0)}" is out of range for type vector')
        vals.append(v)
        i = m.end()
        skip()
        if i < len(s) and s[i] == ",":
            i += 1
        elif i < len(s) and s[i] == "]":
            i += 1
            break
        else:
            raise syntax()
    skip()
    if i != len(s):
        raise syntax()
    if expected_dim is not None and len(vals) != expected_dim:
        raise WireError("dim", f"expected {expected_dim} dimensions, not {len(vals)}")
    return np.array(vals, dtype=np.float32)


def langchain_text(x) -> str:
    """How langchain-postgres 0.0.16 [UPSTREAM] serialises an embedding: str() of the float list."""
    return str([float(v) for v in x])
