"""CPU tests of the files -> jplace pipeline's host pieces (epa-ng_b200/csrc/host/fastio.cpp):
the printf-exact fixed-point formatter and the memory-mapped, multi-threaded query reader."""
import math
import os
import random
import struct

import numpy as np
import pytest


def test_format_fixed_matches_printf(built):
    fmt = built.session.format_fixed
    rng = random.Random(7)
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, 1.5, 2.5, 0.05, 0.15, 0.25, 0.35, 1e-10, 5e-11, 4.9999999999e-11, 1e-300, 5e-324,
            0.1053605157, -5031.3339285153, 0.99999, 1.0 - 2 ** -53, 123456789.987654321, 999999999999999.0, 1e15, 1.7e308,
            2 ** 52 + 0.5, 0.000000000049999999999999, 0.00000000005, 8.5, 9.5, 0.125, 0.375]
    for _ in range(4000):
        vals.append(rng.uniform(-1e4, 1e4))
        vals.append(rng.uniform(0, 1))
        vals.append(math.ldexp(rng.random(), rng.randint(-80, 45)))
        # exact halves of the last printed digit
        vals.append(rng.randint(0, 10 ** 6) / 8.0)
        vals.append(struct.unpack("<d", struct.pack("<Q", rng.getrandbits(62)))[0])
    for p in (0, 1, 3, 6, 10, 15, 18):
        for v in vals:
            assert fmt(v, p) == "%.*f" % (p, v), (v, p)


def _write(path, text):
    with open(path, "w", newline="") as fh:
        fh.write(text)


def test_parallel_reader_equals_serial_reader(built, tmp_path):
    rng = np.random.default_rng(3)
    n, sites = 777, 403
    alphabet = np.frombuffer(b"ACGTacgtNn-?.RYKMSWBDHVOX", dtype=np.uint8)
    rows = alphabet[rng.integers(0, len(alphabet), size=(n, sites))]
    rows[:, 17] = ord("-")                  # all-gap columns
    rows[:, 200] = ord("n")
    rows[:, 402] = ord("?")
    names = ["q%05d some description \"quoted\"" % i for i in range(n)]
    # mixed line wrapping, CRLF, trailing blanks, blank lines
    parts = []
    for i in range(n):
        seq = rows[i].tobytes().decode()
        parts.append(">" + names[i] + (" \t" if i % 5 == 0 else "") + ("\r\n" if i % 3 == 0 else "\n"))
        if i % 4 == 0:
            parts.append(seq + "\n")
        else:
            wrap = 60 + (i % 7)
            for k in range(0, sites, wrap):
                parts.append(seq[k:k + wrap] + ("\r\n" if i % 3 == 0 else "\n"))
        if i % 11 == 0:
            parts.append("\n")
    path = str(tmp_path / "q.fasta")
    _write(path, "".join(parts))
    names_s, rows_s = built.session.read_alignment(path)
    for threads in (1, 3, 8):
        names_p, rows_p, mask = built.session.read_alignment_mt(path, threads)
        assert names_p == names_s == names
        assert np.array_equal(rows_p, rows_s)
        want = np.all(np.isin(rows_s, np.frombuffer(b"NOX.-?", dtype=np.uint8)), axis=0)
        assert np.array_equal(mask.astype(bool), want)
        assert mask[17] and mask[200] and mask[402]
    # the same file as bfast (DNA codes only)
    dna = np.frombuffer(b"-TGKCYSBAWRDMHVN", dtype=np.uint8)[rng.integers(0, 16, size=(50, 33))]
    p2 = str(tmp_path / "d.fasta")
    _write(p2, "".join(">s%d\n%s\n" % (i, dna[i].tobytes().decode()) for i in range(50)))
    bf = built.session.fasta_to_bfast(p2, str(tmp_path))
    nb, rb, mb = built.session.read_alignment_mt(bf, 4)
    assert nb == ["s%d" % i for i in range(50)] and np.array_equal(rb, dna)
    assert np.array_equal(mb.astype(bool), np.all(np.isin(dna, np.frombuffer(b"N-", dtype=np.uint8)), axis=0))


def test_parallel_reader_errors(built, tmp_path):
    p = str(tmp_path / "bad.fasta")
    _write(p, ">a\nACGT\n>b\nACG\n")
    with pytest.raises(built.capi.EpaError, match="equal size"):
        built.session.read_alignment_mt(p, 2)
    _write(p, "ACGT\n>a\nACGT\n")
    with pytest.raises(built.capi.EpaError, match="before the first"):
        built.session.read_alignment_mt(p, 2)
    _write(p, "\n\n")
    with pytest.raises(built.capi.EpaError, match="no sequences"):
        built.session.read_alignment_mt(p, 2)
    with pytest.raises(built.capi.EpaError, match="Cannot open"):
        built.session.read_alignment_mt(str(tmp_path / "missing.fasta"), 2)


def test_bfast_index_is_parallel_and_checks_the_offset_table(built, tmp_path):
    """The bfast records are parsed by all threads from the file's random-access table (Binary_Fasta.hpp:53-64): the same
    rows for any thread count, odd widths included; a table that does not match the entries, or a cut file, is an error."""
    rng = np.random.default_rng(11)
    for sites in (33, 64):
        dna = np.frombuffer(b"-TGKCYSBAWRDMHVN", dtype=np.uint8)[rng.integers(0, 16, size=(301, sites))]
        fa = str(tmp_path / ("w%d.fasta" % sites))
        _write(fa, "".join(">name_%d_%s\n%s\n" % (i, "x" * (i % 9), dna[i].tobytes().decode()) for i in range(301)))
        bf = built.session.fasta_to_bfast(fa, str(tmp_path))
        for threads in (1, 2, 7):
            names, rows, _ = built.session.read_alignment_mt(bf, threads)
            assert names == ["name_%d_%s" % (i, "x" * (i % 9)) for i in range(301)] and np.array_equal(rows, dna)
    raw = bytearray(open(bf, "rb").read())
    table = 7 + 8 + 8 + 64                                  # magic, count, mask length, mask
    bad = bytearray(raw)
    bad[table + 16 * 5 + 8] ^= 0x04                         # offset of entry 5
    p_bad = str(tmp_path / "bad.bfast")
    open(p_bad, "wb").write(bad)
    with pytest.raises(built.capi.EpaError, match="offset table|truncated|equal size"):
        built.session.read_alignment_mt(p_bad, 3)
    p_cut = str(tmp_path / "cut.bfast")
    open(p_cut, "wb").write(raw[: len(raw) - 40])
    with pytest.raises(built.capi.EpaError, match="truncated"):
        built.session.read_alignment_mt(p_cut, 3)


def test_white_space_inside_sequence_lines(built, tmp_path):
    """Blanks and tabs inside sequence lines, lower case and lines of every length around the 16-byte steps of the
    vectorised paths: the parallel reader (white-space count, upper-casing copy) equals the serial one."""
    rng = np.random.default_rng(9)
    n, sites = 211, 131
    alphabet = np.frombuffer(b"ACGTacgtNn-RYKM", dtype=np.uint8)
    rows = alphabet[rng.integers(0, len(alphabet), size=(n, sites))]
    parts = []
    for i in range(n):
        seq = rows[i].tobytes().decode()
        parts.append(">s%d\n" % i)
        if i % 3 == 0:
            parts.append(seq + "\n")                                   # one plain line (fast path)
        elif i % 3 == 1:
            cut = 1 + (i % 40)
            parts.append(seq[:cut] + " \t " + seq[cut:] + "  \n")     # blanks inside and at the end of the line
        else:
            w = 15 + (i % 5)                                           # 15..19 characters per line
            parts.append("".join(seq[k:k + w] + "\n" for k in range(0, sites, w)))
    path = str(tmp_path / "ws.fasta")
    _write(path, "".join(parts))
    names_s, rows_s = built.session.read_alignment(path)
    assert np.array_equal(rows_s, np.char.upper(rows.view("S1")).view(np.uint8).reshape(n, sites))
    for threads in (1, 4):
        names_p, rows_p, _ = built.session.read_alignment_mt(path, threads)
        assert names_p == names_s and np.array_equal(rows_p, rows_s)
