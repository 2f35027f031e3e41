"""pgvector wire formats (SURVEY.md 8f-1 cold-start load, 8a3 text query vector).

CPU part: the oracle restatement (oracle/pgvector_wire.py) against hand-built known answers, the host
parsers of liborx.so (`orx_parse_vector_text`, the dry-run `orx_pgcopy_*` framing walker) against the
oracle.  GPU part: a COPY BINARY stream fed in ragged chunks must leave exactly the rows the oracle
decodes in the table (bit for bit), with pgvector's rejections."""
import struct

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import outline_rag_b200 as orx
from oracle import cosine_topk as O
from oracle import pgvector_wire as W

DIM = 1024


def _rows(n, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, DIM)).astype(np.float32)
    ids = O.ids_from_ints([int(v) for v in rng.integers(1, 2**62, size=n)] if n else [])
    return ids, X


def _feed_all(ld, stream, step):
    for i in range(0, len(stream), step):
        ld.feed(stream[i:i + step])


# ------------------------------------------------------------------------------- oracle known answers
def test_vector_send_known_bytes():
    # dim=2, unused=0, 1.0 = 0x3F800000, -2.5 = 0xC0200000, big-endian
    assert W.vector_send(np.array([1.0, -2.5], np.float32)) == bytes.fromhex("00020000" "3f800000" "c0200000")
    assert np.array_equal(W.vector_recv(bytes.fromhex("00020000" "3f800000" "c0200000")), np.array([1.0, -2.5], np.float32))


def test_vector_recv_rejections():
    good = W.vector_send(np.ones(3, np.float32))
    with pytest.raises(W.WireError, match="unused"):
        W.vector_recv(good[:2] + b"\x00\x01" + good[4:])
    with pytest.raises(W.WireError, match="at least 1"):
        W.vector_recv(struct.pack(">hh", 0, 0))
    with pytest.raises(W.WireError, match="expected 1024"):
        W.vector_recv(good, 1024)
    with pytest.raises(W.WireError, match="NaN"):
        W.vector_recv(struct.pack(">hh", 1, 0) + bytes.fromhex("7fc00000"))
    with pytest.raises(W.WireError, match="infinite"):
        W.vector_recv(struct.pack(">hh", 1, 0) + bytes.fromhex("ff800000"))


def test_copy_stream_known_layout():
    ids = O.ids_from_ints([0x0102030405060708090A0B0C0D0E0F10])
    s = W.copy_binary_stream(ids, np.array([[0.5]], np.float32))
    assert s[:11] == b"PGCOPY\n\xff\r\n\x00" and s[11:19] == bytes(8)
    assert s[19:21] == b"\x00\x02" and s[21:25] == b"\x00\x00\x00\x10"
    assert s[25:41] == bytes(range(1, 17))                       # uuid_send: most significant byte first
    assert s[41:45] == b"\x00\x00\x00\x08" and s[45:49] == b"\x00\x01\x00\x00" and s[49:53] == bytes.fromhex("3f000000")
    assert s[53:] == b"\xff\xff"
    i2, X2, nn = W.copy_binary_parse(s, expected_dim=1)
    assert O.ids_to_ints(i2) == O.ids_to_ints(ids) and X2[0, 0] == 0.5 and nn == 0


def test_strtof_rounding_of_the_text_oracle():
    # 1 + 2^-24 is the midpoint of 1.0 and nextafter(1.0): ties go to even (1.0); a hair above rounds up.
    assert W.vector_in("[1.000000059604644775390625]")[0] == np.float32(1.0)
    assert W.vector_in("[1.0000000596046448]")[0] == np.nextafter(np.float32(1.0), np.float32(2.0))
    # ... which is where str(double) -> strtof differs from double -> float32 (round twice)
    d = 1.0 + 2.0 ** -24
    assert np.float32(d) == np.float32(1.0) and W.vector_in(W.langchain_text([d]))[0] != np.float32(d)
    assert W.vector_in("[3.4028235e38]")[0] == np.finfo(np.float32).max
    assert W.vector_in("[1e-46]")[0] == 0.0 and W.vector_in("[1e-45]")[0] == np.float32(1.4e-45)   # underflow accepted
    with pytest.raises(W.WireError, match="out of range"):
        W.vector_in("[3.5e38]")


# ------------------------------------------------------------------------------- C text parser vs oracle
@pytest.mark.parametrize("text", [
    "[1,2,3]", " [ 1 , 2.5 ,\t-3e-3 ]\n", "[0.1, 0.2, 0.30000000000000004]", "[1e10, -1E-10, +.5]", "[5., 0, -0]",
    "[1.000000059604644775390625, 1.0000000596046448, 16777217]", "[1e-46, 1e-45, 3.4028235e38]",
])
def test_parse_vector_text_matches_the_oracle(text):
    want = W.vector_in(text)
    got = orx.parse_vector_text(text, dim=want.shape[0])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("text,code,msg", [
    ("1,2,3", orx._lib.ORX_ERR_INVALID, "invalid input syntax"),
    ("[]", orx._lib.ORX_ERR_DIM, "at least 1 dimension"),
    ("[1,2,3", orx._lib.ORX_ERR_INVALID, "invalid input syntax"),
    ("[1,,3]", orx._lib.ORX_ERR_INVALID, "invalid input syntax"),
    ("[1,2,3] x", orx._lib.ORX_ERR_INVALID, "Junk after closing"),
    ("[1,a,3]", orx._lib.ORX_ERR_INVALID, "invalid input syntax"),
    ("[1,nan,3]", orx._lib.ORX_ERR_NONFINITE, "NaN not allowed"),
    ("[1,-inf,3]", orx._lib.ORX_ERR_NONFINITE, "infinite value not allowed"),
    ("[1,1e39,3]", orx._lib.ORX_ERR_INVALID, "out of range"),
    ("[1,2]", orx._lib.ORX_ERR_DIM, "expected 3 dimensions, not 2"),
    ("[1,2,3,4]", orx._lib.ORX_ERR_DIM, "expected 3 dimensions, not 4"),
])
def test_parse_vector_text_rejections(text, code, msg):
    with pytest.raises(orx.OrxValueError, match=msg) as ei:
        orx.parse_vector_text(text, dim=3)
    assert ei.value.code == code
    with pytest.raises(W.WireError):
        W.vector_in(text, expected_dim=3)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2**31 - 1))
def test_parse_vector_text_langchain_serialisation(seed):
    """str(list of Python floats) of an fp32 embedding parses back to the same fp32 bits; of arbitrary
    doubles it parses to the correctly rounded fp32 (oracle: exact rational arithmetic)."""
    rng = np.random.default_rng(seed)
    x32 = (rng.standard_normal(DIM) * 10.0 ** rng.integers(-6, 6)).astype(np.float32)
    got = orx.parse_vector_text(W.langchain_text(x32))
    assert np.array_equal(got.view(np.uint32), x32.view(np.uint32))
    x64 = rng.standard_normal(64) * 10.0 ** rng.integers(-30, 30, size=64).astype(np.float64)
    x64 = x64[np.abs(x64) < 3e38]
    text = W.langchain_text(x64)
    got = orx.parse_vector_text(text, dim=x64.shape[0])
    assert np.array_equal(got.view(np.uint32), W.vector_in(text).view(np.uint32))


@settings(max_examples=300, deadline=None)
@given(st.text(alphabet="[], \t0123456789.eE+-nNaAiIfF", min_size=0, max_size=24), st.booleans())
def test_fuzzed_vector_text_is_accepted_or_rejected_exactly_like_the_oracle(body, wrap):
    """Random token soup (decimal syntax only: hex floats, which strtof also takes, are outside the oracle):
    same accept / reject decision as the restatement of vector_in, same fp32 bits when accepted."""
    text = f"[{body}]" if wrap else body
    try:
        want = W.vector_in(text)
    except W.WireError:
        want = None
    try:
        got = orx.parse_vector_text(text, dim=want.shape[0] if want is not None else 3)
    except orx.OrxValueError:
        got = None
    assert (got is None) == (want is None), (text, want, got)
    if want is not None:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), text


# ------------------------------------------------------------------------------- COPY framing, dry run (host)
@pytest.mark.parametrize("step", [1, 7, 4126, 4127, 65536, 1 << 30])
def test_dry_run_loader_counts_rows_for_any_chunking(step):
    n = 40 if step > 1 else 6
    ids, X = _rows(n, seed=step % 97)
    stream = W.copy_binary_stream(ids, X, null_rows=[1, n - 1], header_extension=b"ext-bytes")
    ld = orx.PgCopyLoader(None)
    _feed_all(ld, stream, step)
    assert ld.close() == (n - 2, 2)
    _, X2, nn = W.copy_binary_parse(stream)
    assert X2.shape[0] == n - 2 and nn == 2


def test_dry_run_loader_accepts_eof_at_a_tuple_boundary_and_empty_tables():
    ids, X = _rows(3)
    ld = orx.PgCopyLoader(None)
    ld.feed(W.copy_binary_stream(ids, X, trailer=False))
    assert ld.close() == (3, 0)
    ld = orx.PgCopyLoader(None)
    ld.feed(W.copy_binary_stream(ids[:0], X[:0]))
    assert ld.close() == (0, 0)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_dry_run_keeps_exactly_the_rows_shard_of_assigns(world):
    """Row-sharded cold start: every rank reads the whole stream and keeps mix64(id) mod G == rank -- the C
    side must agree with `sharded.shard_of`, bit for bit, or rows would be lost / duplicated across GPUs."""
    from outline_rag_b200.sharded import shard_of
    rng = np.random.default_rng(world)
    n = 600
    ids = np.stack([rng.integers(0, 2**63, size=n).astype(np.uint64) * np.uint64(2) + np.uint64(1),
                    rng.integers(0, 2**63, size=n).astype(np.uint64)], axis=1)
    ids[:50, 0] = 0                                                # small ids (uuid.UUID(int=i)) too
    X = rng.standard_normal((n, DIM)).astype(np.float32)
    stream = W.copy_binary_stream(ids, X, null_rows=[7, 8])
    live = np.ones(n, bool)
    live[[7, 8]] = False
    owner = shard_of(ids, world)
    total = 0
    for rank in range(world):
        ld = orx.PgCopyLoader(None, world, rank)
        _feed_all(ld, stream, 50_000)
        rows, nulls = ld.close()
        assert rows == int(((owner == rank) & live).sum()) and nulls == 2
        total += rows
    assert total == n - 2
    with pytest.raises(orx.OrxValueError, match="bad world/rank"):
        orx.PgCopyLoader(None, 2, 2)


def _corrupt(stream: bytes, at: int, repl: bytes) -> bytes:
    return stream[:at] + repl + stream[at + len(repl):]


@pytest.mark.parametrize("name,mutate,code,msg", [
    ("signature", lambda s: _corrupt(s, 0, b"XGCOPY"), orx._lib.ORX_ERR_INVALID, "signature not recognized"),
    ("oids", lambda s: _corrupt(s, 11, struct.pack(">i", 1 << 16)), orx._lib.ORX_ERR_INVALID, "WITH OIDS"),
    ("critical", lambda s: _corrupt(s, 11, struct.pack(">i", 1 << 20)), orx._lib.ORX_ERR_INVALID, "critical flags"),
    ("columns", lambda s: _corrupt(s, 19, struct.pack(">h", 3)), orx._lib.ORX_ERR_INVALID, "3 columns"),
    ("null id", lambda s: _corrupt(s, 21, struct.pack(">i", -1)), orx._lib.ORX_ERR_INVALID, "langchain_id"),
    ("dim", lambda s: _corrupt(s, 45, struct.pack(">h", 768)), orx._lib.ORX_ERR_DIM, "expected 1024 dimensions, not 768"),
    ("unused", lambda s: _corrupt(s, 47, struct.pack(">h", 5)), orx._lib.ORX_ERR_INVALID, "unused to be 0, not 5"),
    ("length", lambda s: _corrupt(s, 41, struct.pack(">i", 4096)), orx._lib.ORX_ERR_INVALID, "vector of 4096 bytes"),
    ("nan", lambda s: _corrupt(s, 49 + 4 * 17, bytes.fromhex("7fc00001")), orx._lib.ORX_ERR_NONFINITE, "NaN not allowed"),
    ("inf", lambda s: _corrupt(s, 49 + 4 * 1023, bytes.fromhex("7f800000")), orx._lib.ORX_ERR_NONFINITE, "infinite value"),
    ("after eof", lambda s: s + b"\x00", orx._lib.ORX_ERR_INVALID, "after EOF marker"),
    ("truncated", lambda s: s[:-2 - 100], orx._lib.ORX_ERR_INVALID, "unexpected EOF"),
    ("short header", lambda s: s[:10], orx._lib.ORX_ERR_INVALID, "invalid COPY file header"),
])
def test_dry_run_loader_rejects_what_postgres_and_pgvector_reject(name, mutate, code, msg):
    ids, X = _rows(2, seed=3)
    bad = mutate(W.copy_binary_stream(ids, X))
    with pytest.raises(orx.OrxError, match=msg) as ei:
        ld = orx.PgCopyLoader(None)
        _feed_all(ld, bad, 1000)
        ld.close()
    assert ei.value.code == code
    with pytest.raises(W.WireError):
        W.copy_binary_parse(bad)


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 2**31 - 1))
def test_fuzzed_streams_are_accepted_or_rejected_exactly_like_the_oracle(seed):
    """Random byte mutations, truncations and splices of a valid stream, fed in random chunk sizes: the C
    walker never crashes, accepts exactly what the strict Python restatement accepts, and counts the same rows."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 5))
    ids, X = _rows(n, seed=seed % 1000)
    nulls = [int(i) for i in np.nonzero(rng.random(n) < 0.3)[0]]
    s = bytearray(W.copy_binary_stream(ids, X, null_rows=nulls, header_extension=bytes(int(rng.integers(0, 6))),
                                       trailer=bool(rng.random() < 0.8)))
    kind = int(rng.integers(0, 5))
    if kind == 0:                                                   # corrupt framing bytes near the front
        for _ in range(int(rng.integers(1, 4))):
            s[int(rng.integers(0, min(len(s), 80)))] = int(rng.integers(0, 256))
    elif kind == 1:                                                 # corrupt bytes anywhere (mostly payload)
        for _ in range(int(rng.integers(1, 4))):
            s[int(rng.integers(0, len(s)))] = int(rng.integers(0, 256))
    elif kind == 2:                                                 # truncate
        del s[int(rng.integers(0, len(s))):]
    elif kind == 3:                                                 # splice garbage in / append after the trailer
        at = int(rng.integers(0, len(s) + 1))
        s[at:at] = rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8).tobytes()
    s = bytes(s)
    try:
        _, Xo, nn = W.copy_binary_parse(s)
        want = (Xo.shape[0], nn)
    except W.WireError:
        want = None
    got = None
    ld = orx.PgCopyLoader(None)
    try:
        p = 0
        while p < len(s):
            step = int(rng.integers(1, 6000))
            ld.feed(s[p:p + step])
            p += step
        got = ld.close()
    except orx.OrxError:
        try:
            ld.close()
        except orx.OrxError:
            pass
    assert got == want, (kind, want, got)


def test_loader_fails_for_good_after_an_error():
    ids, X = _rows(2)
    ld = orx.PgCopyLoader(None)
    with pytest.raises(orx.OrxError):
        ld.feed(b"not a copy stream, but long enough")
    with pytest.raises(orx.OrxError, match="already failed"):
        ld.feed(W.copy_binary_stream(ids, X))
    with pytest.raises(orx.OrxError):
        ld.close()
    assert ld.close() == (0, 0)        # closing twice is harmless


def test_close_repeats_the_loaders_own_error_even_from_another_thread():
    """`aload_pgcopy` feeds and closes on worker threads (`asyncio.to_thread`): the error text is kept by the loader,
    not in the feeding thread's last-error slot."""
    import threading
    ld = orx.PgCopyLoader(None)
    with pytest.raises(orx.OrxValueError, match="signature not recognized"):
        ld.feed(b"not a copy stream, but long enough")
    seen = []

    def close_elsewhere():
        try:
            ld.close()
        except orx.OrxError as e:
            seen.append(str(e))

    t = threading.Thread(target=close_elsewhere)
    t.start()
    t.join()
    assert len(seen) == 1 and "signature not recognized" in seen[0]


# ------------------------------------------------------------------------------- GPU: the load itself
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("step", [3001, 1 << 20, 1 << 31])
def test_pgcopy_load_leaves_the_oracles_rows_in_the_table(dtype, step):
    from tests._helpers import stored_bf16_rows
    n = 700
    ids, X = _rows(n, seed=11)
    X[5, 100] = np.float32(1e-42)                                   # a denormal survives the byte swap
    X[6] = 0.0                                                      # a zero row is legal (distance NaN)
    nulls = [0, 350, n - 1]
    stream = W.copy_binary_stream(ids, X, null_rows=nulls, header_extension=b"x" * 5)   # odd payload alignment
    want_ids, want_X, want_null = W.copy_binary_parse(stream)
    with orx.Index(dtype) as ix:
        assert ix.load_pgcopy(stream[i:i + step] for i in range(0, len(stream), step)) == (n - 3, want_null)
        assert len(ix) == n - 3
        got, found = ix.fetch(want_ids)
        assert found.all()
        if dtype == "fp32":
            assert np.array_equal(got.view(np.uint32), want_X.view(np.uint32))          # bit for bit
        else:
            keep = np.ones(n - 3, bool)
            keep[np.where((want_X == 0).all(axis=1))[0]] = False
            assert np.array_equal(got[keep], stored_bf16_rows(want_X[keep]))
        assert not ix.fetch(ids[nulls])[1].any()
        # and the loaded table answers like one built by upsert
        q = X[17] + 0.1 * X[18]
        g_ids, g_d, _ = ix.search(q, 12)
        with orx.Index(dtype) as ref:
            ref.upsert(want_ids, want_X)
            r_ids, r_d, _ = ref.search(q, 12)
        assert np.array_equal(g_ids, r_ids) and np.array_equal(g_d.view(np.uint64), r_d.view(np.uint64))


@pytest.mark.gpu
def test_pgcopy_load_spans_several_flushes_and_upserts_repeated_ids():
    n = 16384 + 16384 + 100                                         # three device batches
    rng = np.random.default_rng(5)
    base = rng.standard_normal((64, DIM)).astype(np.float32)
    X = base[rng.integers(0, 64, size=n)] * rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)
    ids = O.ids_from_ints(list(range(1, n + 1)))
    ids[n - 1] = ids[3]                                             # the stream repeats an id: last one wins
    stream = W.copy_binary_stream(ids, X)
    with orx.Index("fp32") as ix:
        with ix.pgcopy_loader() as ld:
            _feed_all(ld, stream, 10_000_019)
        assert ld.result == (n, 0) and len(ix) == n - 1
        probe = np.array([0, 3, 16383, 16384, 32767, 32768, n - 2])
        got, found = ix.fetch(ids[probe])
        want = X[probe].copy()
        want[1] = X[n - 1]
        assert found.all() and np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_pgcopy_load_rejects_a_nan_batch_and_keeps_earlier_batches():
    n = 16384 + 10
    ids, X = _rows(n, seed=2)
    X[16384 + 4, 9] = np.nan                                        # in the second device batch
    stream = W.copy_binary_stream(ids, X)
    with orx.Index("fp32") as ix:
        ld = ix.pgcopy_loader()
        ld.feed(stream[:5_000_000])
        with pytest.raises(orx.OrxValueError, match="NaN or infinite") as ei:
            ld.feed(stream[5_000_000:])
            ld.close()
        assert ei.value.code == orx._lib.ORX_ERR_NONFINITE
        assert ld.result == (16384, 0) and len(ix) == 16384        # the batch before the bad one stays loaded
    with orx.Index("fp32") as ix, pytest.raises(orx.OrxValueError, match="expected 1024 dimensions"):
        ix.load_pgcopy(W.copy_binary_stream(ids[:2], X[:2, :512]))


@pytest.mark.gpu
def test_parallel_copy_streams_load_one_index_while_it_is_searched():
    """Cold start over several connections (each COPY selects a slice of the ids): one loader per thread, each with its
    own staging buffers and stream, all committing into ONE index; a reader thread searches meanwhile.  The table ends up
    with exactly the rows of all streams and answers like the oracle."""
    import threading
    per, n_streams = 16384 * 2 + 77, 3                              # every loader flushes three device batches
    parts = [_rows(per, seed=10 + i) for i in range(n_streams)]
    for i, (ids_i, _) in enumerate(parts):
        ids_i[:, 1] += np.uint64(i * 1_000_000)                     # disjoint id ranges
    streams = [W.copy_binary_stream(ids_i, X_i) for ids_i, X_i in parts]
    errors, results = [], [None] * n_streams
    with orx.Index("fp32") as ix:
        def load(i):
            try:
                with ix.pgcopy_loader() as ld:
                    _feed_all(ld, streams[i], 1_000_003 + 17 * i)
                results[i] = ld.result
            except Exception as e:      # noqa: BLE001
                errors.append(e)

        stop = threading.Event()

        def read():
            try:
                while not stop.is_set():
                    if len(ix):
                        got = ix.search(parts[0][1][:3], 4)
                        assert (got[2] >= 1).all()
            except Exception as e:      # noqa: BLE001
                errors.append(e)

        th = [threading.Thread(target=load, args=(i,)) for i in range(n_streams)] + [threading.Thread(target=read)]
        for t in th:
            t.start()
        for t in th[:-1]:
            t.join()
        stop.set()
        th[-1].join()
        assert not errors, errors
        assert results == [(per, 0)] * n_streams and len(ix) == per * n_streams
        all_ids = np.concatenate([p[0] for p in parts])
        all_X = np.concatenate([p[1] for p in parts])
        probe = np.array([0, per - 1, per, 2 * per + 5, 3 * per - 1])
        got, found = ix.fetch(all_ids[probe])
        assert found.all() and np.array_equal(got.view(np.uint32), all_X[probe].view(np.uint32))
        Q = all_X[[5, per + 9]] + np.float32(0.01)
        g_ids, g_d, g_c = ix.search(Q, 12)
        for i in range(2):
            w_ids, w_d = O.topk_exact(all_X, all_ids, Q[i], 12)
            assert np.array_equal(g_ids[i], w_ids) and np.array_equal(g_d[i].view(np.uint64), w_d.view(np.uint64))


@pytest.mark.gpu
def test_closing_the_index_under_an_open_loader_settles_its_batch_in_flight():
    """The loader commits full batches on a helper thread: `Index.close()` waits for that batch and frees the loader
    before the table goes away; the abandoned loader then refuses further input."""
    n = 16384 + 50                                                  # one full batch is handed to the helper thread
    ids, X = _rows(n, seed=4)
    stream = W.copy_binary_stream(ids, X)
    ix = orx.Index("fp32")
    ld = ix.pgcopy_loader()
    ld.feed(stream)
    ix.close()
    with pytest.raises(orx.OrxValueError, match="closed"):
        ld.feed(b"x")
    assert ld.close() == (0, 0)


@pytest.mark.gpu
def test_sharded_pgcopy_load_partitions_the_stream_without_loss():
    from outline_rag_b200.sharded import shard_of
    n = 900
    ids, X = _rows(n, seed=13)
    stream = W.copy_binary_stream(ids, X, null_rows=[5])
    owner = shard_of(ids, 2)
    owner[5] = -1
    with orx.Index("fp32") as a, orx.Index("fp32") as b:
        ra = a.load_pgcopy([stream[:300_000], stream[300_000:]], world=2, rank=0)
        rb = b.load_pgcopy(stream, world=2, rank=1)
        assert ra == (int((owner == 0).sum()), 1) and rb == (int((owner == 1).sum()), 1)
        for ix, r in ((a, 0), (b, 1)):
            got, found = ix.fetch(ids)
            assert np.array_equal(found, owner == r)
            assert np.array_equal(got[found].view(np.uint32), X[owner == r].view(np.uint32))
        # the two shards merge to the answer of one table holding every row
        q = X[40] + 0.05 * X[41]
        pa, pb = a.search(q, 12), b.search(q, 12)
        merged = a.merge_topk(np.stack([pa[0], pb[0]]), np.stack([pa[1], pb[1]]), np.stack([pa[2], pb[2]]), 12)
        keep = owner >= 0
        w_ids, w_d = O.topk_exact(X[keep], ids[keep], q, 12)
        assert np.array_equal(merged[0][0], w_ids) and np.array_equal(merged[1][0].view(np.uint64), w_d.view(np.uint64))


@pytest.mark.gpu
def test_vectorstore_cold_start_from_an_async_copy_stream():
    import asyncio

    class Emb:
        async def aembed_query(self, text):
            return [0.0] * DIM

    ids, X = _rows(300, seed=9)
    stream = W.copy_binary_stream(ids, X)

    async def chunks():
        for i in range(0, len(stream), 8192):
            yield memoryview(stream)[i:i + 8192]

    async def run():
        store = await orx.GpuVectorStore.create(None, Emb())
        # content and metadata live in the doc store (Postgres in production); only the vectors are streamed
        sids = orx.ids_to_uuid_strs(ids)
        store.doc_store.put_many(sids, [f"chunk {i}" for i in range(300)], [{"source_id": "d"}] * 300)
        assert await store.aload_pgcopy(chunks(), feed_bytes=100_000) == (300, 0)
        # the text form of the query vector, as the reference sends it
        q = orx.parse_vector_text(W.langchain_text(X[42]))
        hits = await store.asimilarity_search_with_score_by_vector(q, k=3)
        assert hits[0][0].id == sids[42] and hits[0][0].page_content == "chunk 42" and hits[0][1] < 1e-12
        # a chunk that has left the source of truth is not returned (the SQL would not return it either)
        store.doc_store.delete_many([sids[42]])
        assert all(h[0].id != sids[42] for h in await store.asimilarity_search_with_score_by_vector(q, k=3))
        store.index.close()

    asyncio.run(run())


# ------------------------------------------------------------------------------- host-side encoder
def test_encoder_emits_exactly_the_oracles_stream():
    ids, X = _rows(37, seed=6)
    assert orx.encode_copy_binary(ids, X) == W.copy_binary_stream(ids, X)
    from outline_rag_b200.pgwire import COPY_HEADER, COPY_TRAILER, encode_tuples
    parts = COPY_HEADER + encode_tuples(ids[:20], X[:20]) + encode_tuples(ids[20:], X[20:]) + COPY_TRAILER
    assert parts == W.copy_binary_stream(ids, X)
    import uuid
    as_strings = [str(uuid.UUID(int=(int(h) << 64) | int(l))) for h, l in ids]
    assert orx.encode_copy_binary(as_strings, X.tolist()) == W.copy_binary_stream(ids, X)
    ld = orx.PgCopyLoader(None)
    ld.feed(orx.encode_copy_binary(ids, X))
    assert ld.close() == (37, 0)
    with pytest.raises(ValueError):
        orx.encode_copy_binary(ids[:3], X)
