"""Wire formats either side of the path: uos point files, .pose files, .frames files and the
frame rule of Scan::transform.  Host only."""
import os

import numpy as np
import pytest

REF_DAT = "/root/reference/dat"          # present in the build container only; never on the GPU box


def _write(path, text):
    with open(path, "w", newline="") as f:
        f.write(text)


def test_read_uos_plain_and_decorated(icp, tmp_path):
    rng = np.random.default_rng(3)
    pts = np.round(rng.normal(0, 300, (5000, 3)), 4)
    p = tmp_path / "scan000.3d"
    _write(p, "".join("%s %s %s\n" % tuple(repr(float(v)) for v in row) for row in pts))
    assert np.array_equal(icp.read_uos(p), pts)
    # comments, blank lines, tabs, several blanks, CRLF, explicit signs and exponents (handle_line, helper.cc:577-640)
    _write(p, "# header comment\r\n\r\n  1 2\t3\r\n+4.5   -6e1 7.25 # trailing comment\r\n\t\n8 9 10")
    assert np.array_equal(icp.read_uos(p), [[1, 2, 3], [4.5, -60, 7.25], [8, 9, 10]])
    _write(p, "")
    assert icp.read_uos(p).shape == (0, 3)
    with pytest.raises(icp.B200ICPError):
        icp.read_uos(tmp_path / "missing.3d")


def test_read_uos_header_rule(icp, tmp_path):
    p = tmp_path / "s.3d"
    # up to 10 unparsable lines at the top are skipped (helper.cc:752-822) ...
    _write(p, "garbage line\n" * 10 + "1 2 3\n")
    assert np.array_equal(icp.read_uos(p), [[1, 2, 3]])
    # ... the 11th is fatal
    _write(p, "garbage line\n" * 11 + "1 2 3\n")
    with pytest.raises(icp.B200ICPError):
        icp.read_uos(p)
    # once a line has been handled (a point, but also a comment or an empty line) every bad line is fatal
    # ... including out-of-range tokens: strtoval rejects what strtod flags with ERANGE (helper.cc:242-271)
    for text in ("1 2 3\nbad\n", "# c\nbad\n1 2 3\n", "1 2 3\n1 2\n", "1 2 3\n1 2 3 4\n", "1 2 3\n1 2 x3\n",
                 "1 2 3\n1e400 2 3\n", "1 2 3\n1 -1e400 3\n", "1 2 3\n1 2 1e-400\n"):
        _write(p, text)
        with pytest.raises(icp.B200ICPError):
            icp.read_uos(p)
    # wrong arity inside the header window is just one of the tolerated lines
    _write(p, "1 2\n1 2 3 4\n5 6 7\n")
    assert np.array_equal(icp.read_uos(p), [[5, 6, 7]])


def test_read_uos_large_file_parallel_chunks(icp, tmp_path):
    rng = np.random.default_rng(4)
    pts = rng.normal(0, 1000, (120000, 3))
    p = tmp_path / "big.3d"
    _write(p, "".join("%.17g %.17g %.17g\n" % tuple(row) for row in pts))     # > 1 MiB: parsed by several threads
    assert os.path.getsize(p) > (1 << 20)
    assert np.array_equal(icp.read_uos(p), pts)                               # round-trip exact, order preserved
    with open(p, "a") as f:
        f.write("oops\n")
    with pytest.raises(icp.B200ICPError):
        icp.read_uos(p)


def test_read_pose_degrees_to_radians(icp, tmp_path):
    p = tmp_path / "scan001.pose"
    _write(p, "-3.10605 -7.50803 156.917\n1.35694 -0.852409 -0.56224\n")       # the reference's dat/scan001.pose
    pos, th = icp.read_pose(p)
    assert np.array_equal(pos, [-3.10605, -7.50803, 156.917])
    np.testing.assert_allclose(th, np.deg2rad([1.35694, -0.852409, -0.56224]), rtol=1e-15)


@pytest.mark.skipif(not os.path.isdir(REF_DAT), reason="reference dat/ not present")
def test_read_reference_dat_scans(icp):
    for k in range(3):
        got = icp.read_uos(os.path.join(REF_DAT, "scan%03d.3d" % k))
        assert got.shape == (81360, 3)                                         # wc -l dat/scan00*.3d
        if k == 1:
            assert np.array_equal(got, np.loadtxt(os.path.join(REF_DAT, "scan001.3d")))
        pos, th = icp.read_pose(os.path.join(REF_DAT, "scan%03d.pose" % k))
        assert np.all(np.isfinite(pos)) and np.all(np.abs(th) < np.pi)


def test_frame_rule_of_scan_transform(icp):
    I, A, L, X = icp.FRAME_ICPINACTIVE, icp.FRAME_ICP, icp.FRAME_LUM, icp.FRAME_INVALID
    poses = np.stack([np.arange(16.0) + 100 * k for k in range(4)])
    fr = icp.Frames(4)
    fr.transform(2, poses, A, 0)         # islum 0: everybody gets a frame (scan.cc:955-983)
    assert [fr.get(k)[0][1] for k in range(4)] == [I, I, A, X]
    assert all(np.array_equal(fr.get(k)[0][0], poses[k]) for k in range(4))
    fr.transform(0, poses, A, 0)         # `found` stays 0 for scan 0: the others count as inactive
    assert [fr.get(k)[1][1] for k in range(4)] == [A, I, I, I]
    fr.transform(1, poses, A, -1)        # islum -1: nothing
    fr.transform(1, poses, X, 0)         # type INVALID: nothing (scan.cc:941)
    assert [len(fr.get(k)) for k in range(4)] == [2, 2, 2, 2]
    fr.transform(1, poses, L, 1)         # islum 1: only this scan
    assert [len(fr.get(k)) for k in range(4)] == [2, 3, 2, 2] and fr.get(1)[2][1] == L
    fr.transform(2, poses, L, 2)         # islum 2: this scan and scan 0, INVALID for the scans after it
    assert [len(fr.get(k)) for k in range(4)] == [3, 3, 3, 3]
    assert [fr.get(k)[-1][1] for k in range(4)] == [L, L, L, X]
    with pytest.raises(icp.B200ICPError):
        fr.transform(7, poses, A, 0)


def test_frames_file_format(icp, tmp_path):
    fr = icp.Frames(2)
    m = np.array([1, 0, 0, 0, 0, 0.99999994, -1.5e-7, 0, 0, 1.25e-5, 1, 0, -3.10605, 123456789.0, 156.917, 1])
    fr.add(1, m, icp.FRAME_ICP)
    fr.add(1, m * 2, icp.FRAME_LUM)
    p = tmp_path / "scan001.frames"
    fr.save(1, p)
    lines = open(p).read().split("\n")
    assert lines[-1] == "" and len(lines) == 3
    # operator<<(ostream&, const double[16]) (globals.icc:123-132): default precision (%g), blank after EVERY value
    assert lines[0] == "".join("%g " % v for v in m) + "1"
    assert lines[1] == "".join("%g " % v for v in 2 * m) + "3"
    fr.save(1, p, append=True)
    assert len(open(p).read().split("\n")) == 5
    fr.save(0, p)                                     # a scan without frames writes an empty file
    assert open(p).read() == ""
    bad = icp.Frames(1)
    bad.add(0, np.full(16, np.nan), icp.FRAME_ICP)
    with pytest.raises(icp.B200ICPError):
        bad.save(0, p)


# ---- differential test against a line-by-line restatement of readASCII / handle_line (src/scanio/helper.cc:577-835)
def _ref_read_ascii(text):
    """Python restatement (oracle side) of the reference's uos parsing rules.  Returns (points, ok)."""
    pts, header = [], 10
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()                       # getline at EOF after a final newline reads an empty line: handled as success
        trailing_empty = True
    else:
        trailing_empty = False
    for raw in lines + ([""] if trailing_empty else []):
        line = raw[:-1] if raw.endswith("\r") else raw
        s = line.lstrip(" \t")
        good, vals = True, []
        if s != "" and not s.startswith("#"):
            body = s.split("#", 1)[0]
            toks = body.replace("\t", " ").split(" ")
            toks = [t for t in toks if t != ""]
            for t in toks:
                try:
                    if t.lower().lstrip("+-") in ("nan", "inf", "infinity") or t.lower().lstrip("+-").startswith("0x"):
                        raise ValueError      # the generator below never emits these; keep the restatement simple
                    vals.append(float(t))
                except ValueError:
                    good = False
                    break
            if good and len(vals) != 3:
                good = False
            if good:
                pts.append(vals)
        if not good:
            header -= 1
            if header < 0:
                return None, False
        elif header >= 0:
            header = -1
    return np.array(pts, dtype=np.float64).reshape(-1, 3), True


def test_read_uos_differential_random_files(icp, tmp_path):
    rng = np.random.default_rng(11)
    pieces_good = [lambda: "%s %s %s" % tuple(repr(float(v)) for v in rng.normal(0, 100, 3)),
                   lambda: "  %d\t%d   %d  " % tuple(rng.integers(-500, 500, 3)),
                   lambda: "%.3e %.3e %.3e # c" % tuple(rng.normal(0, 1e3, 3)),
                   lambda: "# only a comment", lambda: "", lambda: " \t "]
    pieces_bad = [lambda: "1 2", lambda: "1 2 3 4", lambda: "a b c", lambda: "1 2 3x", lambda: "1,2,3", lambda: "--1 2 3"]
    p = tmp_path / "r.3d"
    outcomes = {True: 0, False: 0}
    for trial in range(300):
        n_head_bad = int(rng.choice([0, 0, 0, 1, 3, 10, 11]))
        lines = [pieces_bad[rng.integers(len(pieces_bad))]() for _ in range(n_head_bad)]
        for _ in range(int(rng.integers(0, 30))):
            lines.append(pieces_good[rng.integers(len(pieces_good))]())
        if rng.random() < 0.15 and lines:
            lines.insert(int(rng.integers(0, len(lines) + 1)), pieces_bad[rng.integers(len(pieces_bad))]())
        eol = "\r\n" if rng.random() < 0.3 else "\n"
        text = eol.join(lines) + (eol if rng.random() < 0.7 else "")
        _write(p, text)
        want, ok = _ref_read_ascii(text)
        outcomes[ok] += 1
        if ok:
            got = icp.read_uos(p)
            assert got.shape == want.shape and np.array_equal(got, want), (trial, text)
        else:
            with pytest.raises(icp.B200ICPError):
                icp.read_uos(p)
    assert outcomes[True] > 100 and outcomes[False] > 10


def test_read_uos_numbers_are_correctly_rounded(icp, tmp_path):
    """the reader's fast decimal path (<= 15 digits, |exp10| <= 22) and its from_chars / strtod fallback must both
    return what strtod returns -- Python's float() is a correctly rounded strtod"""
    rng = np.random.default_rng(23)
    toks = []
    for fmt in ("%.6g", "%.3f", "%.10g", "%.15g", "%.17g", "%.8e", "%d", "%.1f", "%.12f"):
        vals = np.concatenate([rng.normal(0, 1e3, 300), rng.normal(0, 1e-3, 100), rng.uniform(-1e15, 1e15, 100),
                               10.0 ** rng.uniform(-25, 25, 100) * rng.choice([-1, 1], 100)])
        toks += [fmt % (int(v) if fmt == "%d" else v) for v in vals]
    toks += ["0", "-0", "0.0", "-0.000", "1e22", "1e23", "1e-22", "1e-23", "123456789012345", "1234567890123456",
             "0.000000000000001", "9007199254740993", "4.35", "0.1", "2.675", "1.7976931348623157e308", "4.9e-324",
             "5e-324", ".5", "5.", "-.25e1", "1E5", "1e+5", "1e-5", "00012.50", "+3.5"]
    while len(toks) % 3:
        toks.append("1")
    p = tmp_path / "n.3d"
    _write(p, "".join("%s %s %s\n" % tuple(toks[i:i + 3]) for i in range(0, len(toks), 3)))
    got = icp.read_uos(p).reshape(-1)
    want = np.array([float(t) for t in toks])
    assert got.shape == want.shape
    bad = [(t, g, w) for t, g, w in zip(toks, got, want) if not (g == w and np.signbit(g) == np.signbit(w))]
    assert not bad, bad[:5]


def test_frames_load_round_trip_and_rules(icp, tmp_path):
    fr = icp.Frames(2)
    rng = np.random.default_rng(2)
    mats = [rng.normal(0, 10, 16) for _ in range(5)]
    for k, m in enumerate(mats):
        fr.add(0, m, k % 5)
    p = tmp_path / "scan000.frames"
    fr.save(0, p)
    back = icp.Frames(2)
    back.add(1, mats[0], 1)                         # loading replaces the list (readFrames clears it)
    back.load(1, p)
    got = back.get(1)
    assert [t for _, t in got] == [k % 5 for k in range(5)]
    for (m, _), want in zip(got, mats):
        np.testing.assert_allclose(m, want, rtol=1e-5)          # the file holds 6 significant digits
    # empty lines and '#' lines are skipped, anything else malformed is an error (basicScan.cc:882-895)
    _write(p, "# comment\n\n" + " ".join(["1"] * 16) + " 3\n")
    back.load(1, p)
    assert len(back.get(1)) == 1 and back.get(1)[0][1] == 3
    _write(p, " ".join(["1"] * 15) + " x 3\n")
    with pytest.raises(icp.B200ICPError):
        back.load(1, p)
    assert len(back.get(1)) == 1                                  # untouched by the failed load
    with pytest.raises(icp.B200ICPError):
        back.load(1, tmp_path / "missing.frames")


def test_graph_net_file(icp, tmp_path):
    p = tmp_path / "bremen.net"
    _write(p, "4\n3\n0 1\n1 2\n3 4\n4 0\n")                        # the example of graph.cc:36-50: only 3 links are read
    g = icp.Graph.from_net_file(p)
    assert [tuple(l) for l in g.links] == [(0, 1), (1, 2), (3, 4)]
    assert g.n_scans == 5                                         # what addLink counts: ids 0 1 2 3 4
    _write(p, "13 4  0 1  1 2\n2 3 3 0")                          # any whitespace separates
    g = icp.Graph.from_net_file(p)
    assert [tuple(l) for l in g.links] == [(0, 1), (1, 2), (2, 3), (3, 0)] and g.n_scans == 4
    _write(p, "3 5\n0 1\n1 2\n")                                  # promises more links than it holds
    with pytest.raises(icp.B200ICPError):
        icp.Graph.from_net_file(p)
    with pytest.raises(icp.B200ICPError):
        icp.Graph.from_net_file(tmp_path / "missing.net")


def test_write_uos_formats_and_round_trip(icp, tmp_path):
    rng = np.random.default_rng(8)
    pts = np.concatenate([rng.normal(0, 500, (70000, 3)), [[0.0, -0.0, 1e-300], [1e15, -1e-15, 123.456]]])
    p = tmp_path / "scan000.3d"
    icp.write_uos(p, pts, fmt=0)                       # "%lf %lf %lf": 6 decimals
    first = open(p).readline()
    assert first == "%f %f %f\n" % tuple(pts[0])
    assert np.all(np.abs(icp.read_uos(p) - pts) <= 5.1e-7 + 1e-15 * np.abs(pts))
    icp.write_uos(p, pts, fmt=1)                       # "%.016e": round-trips a double
    assert open(p).readline() == "%.016e %.016e %.016e\n" % tuple(pts[0])
    assert np.array_equal(icp.read_uos(p), pts)
    icp.write_uos(p, pts, fmt=2)                       # "%.013a": hex floats, through the reader's strtod fallback
    assert open(p).readline() == "%s %s %s\n" % tuple(_hex13(v) for v in pts[0])
    back = icp.read_uos(p)
    assert np.array_equal(back, pts) and np.array_equal(np.signbit(back), np.signbit(pts))
    icp.write_uos(p, pts[:3], scale=100.0, fmt=1)      # scaleFac
    assert np.array_equal(icp.read_uos(p), 100.0 * pts[:3])
    icp.write_uos(p, np.zeros((0, 3)))
    assert open(p).read() == ""
    with pytest.raises(icp.B200ICPError):
        icp.write_uos(tmp_path / "nodir" / "x.3d", pts[:1])


def _hex13(v):
    """C's %.013a for a double"""
    import ctypes
    buf = ctypes.create_string_buffer(64)
    ctypes.CDLL(None).snprintf(buf, 64, b"%.013a", ctypes.c_double(float(v)))
    return buf.value.decode()
