"""host/fq2psmcfa (consensus FASTQ -> .psmcfa) and host/splitfa against the UNMODIFIED reference utilities.

Golden outputs under tests/golden/fq/ were printed by oracle/_ref/fq2psmcfa (utils/fq2psmcfa.c compiled as it lies in the
reference tree; tools/make_golden_fq.py) for every option the utility has; the bar is byte identity.  Where the reference
binary is present (this container) random inputs are also run through both."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden", "fq")
MINE = os.path.join(ROOT, "host", "fq2psmcfa")
REF = os.path.join(ROOT, "oracle", "_ref", "fq2psmcfa")
ALL = json.load(open(os.path.join(G, "index.json")))
INDEX = {k: v for k, v in ALL.items() if v.get("tool") != "splitfa"}
SPLIT = {k: v for k, v in ALL.items() if v.get("tool") == "splitfa"}
MINE_SPLIT = os.path.join(ROOT, "host", "splitfa")
REF_SPLIT = os.path.join(ROOT, "oracle", "_ref", "splitfa")


def run(binary, args, inp=None, stdin=None):
    r = subprocess.run([binary] + args + ([inp] if inp else []), input=stdin, capture_output=True)
    return r.returncode, r.stdout, r.stderr


@pytest.mark.parametrize("golden", sorted(INDEX))
def test_matches_reference_output_byte_for_byte(golden):
    meta = INDEX[golden]
    want = gzip.open(os.path.join(G, golden), "rb").read()
    rc, got, err = run(MINE, meta["args"], os.path.join(G, meta["input"]))
    assert rc == 0, err
    assert got == want, (golden, len(got), len(want))
    assert got.count(b">") == meta["records"]


@pytest.mark.parametrize("threads", ["2", "5"])
def test_threads_do_not_change_the_output(threads):
    meta = INDEX["cons_fq.g300.psmcfa.gz"]
    want = gzip.open(os.path.join(G, "cons_fq.g300.psmcfa.gz"), "rb").read()
    rc, got, err = run(MINE, meta["args"] + ["-p", threads], os.path.join(G, meta["input"]))
    assert rc == 0 and got == want


def test_stdin_and_gzip_input():
    want = gzip.open(os.path.join(G, "cons_fq.g300.psmcfa.gz"), "rb").read()
    raw = gzip.open(os.path.join(G, "cons.fq.gz"), "rb").read()
    rc, got, _ = run(MINE, ["-g", "300", "-"], stdin=raw)                      # plain text on stdin
    assert rc == 0 and got == want
    rc, got, _ = run(MINE, ["-g", "300", "-"], stdin=gzip.compress(raw))       # gzip on stdin
    assert rc == 0 and got == want


def test_option_errors_like_the_reference():
    rc, out, err = run(MINE, [])
    assert rc == 1 and out == b"" and b"Usage: fq2psmcfa" in err
    rc, out, err = run(MINE, ["-v", "-n", os.path.join(G, "cons.fa")])
    assert rc == 2 and b"only one of the options -c, -n, -v and -C" in err


def test_output_feeds_the_psmcfa_reader(tmp_path):
    """the produced text is what the E-step's input path expects: T/K/N in 60 columns -> 0/1/2 bins (cli.c:103-138)"""
    from psmc_b200 import psmcfa
    rc, got, _ = run(MINE, ["-g", "300"], os.path.join(G, "cons.fq.gz"))
    p = tmp_path / "x.psmcfa"
    p.write_bytes(got)
    names, seqs = psmcfa.read_psmcfa(str(p))
    assert len(seqs) == 7 and names[0] == "chr1"
    body = [l for l in got.decode().split(">")[1].splitlines()[1:]]
    assert all(len(l) == 60 for l in body[:-1]) and set("".join(body)) <= set("TKN")
    assert len(seqs[0]) == len("".join(body)) == (30050 + 99) // 100
    assert int((seqs[0] == 1).sum()) == "".join(body).count("K") and int((seqs[0] == 2).sum()) == "".join(body).count("N")


def _random_input(rng, kind):
    alphabet = list("ACGTACGTACGTMRWSYKVHDBNXacgtnmrwsyk-*.") + ["\t"]
    recs = []
    for r in range(int(rng.integers(1, 6))):
        n = int(rng.integers(0, 4000))
        s = "".join(rng.choice(alphabet, size=n)).replace("\t", "")
        width = int(rng.integers(20, 200))
        lines = [s[i:i + width] for i in range(0, len(s), width)] or [""]
        if kind == "fq":
            q = "".join(chr(int(v)) for v in rng.integers(33, 90, size=len(s)))
            qlines = [q[i:i + width] for i in range(0, len(q), width)] or [""]
            recs.append("@r%d c\n%s\n+r%d\n%s\n" % (r, "\n".join(lines), r, "\n".join(qlines)))
        else:
            recs.append(">r%d\n%s\n" % (r, "\n".join(lines)))
    return "".join(recs).encode()


@pytest.mark.skipif(not os.path.exists(REF), reason="reference utility not built here (oracle/_ref/fq2psmcfa)")
def test_random_inputs_against_the_reference_binary(tmp_path):
    rng = np.random.default_rng(5)
    opts = [[], ["-q", "30"], ["-v"], ["-n"], ["-c"], ["-C"], ["-s", "7"], ["-s", "250", "-q", "45"], ["-x"]]
    for trial in range(40):
        kind = "fq" if trial % 2 == 0 else "fa"
        data = _random_input(rng, kind)
        if kind == "fq":   # qualities may contain '@' and '>' at line starts: only a problem for readers that look for them there, not for kseq's
            pass
        p = tmp_path / ("t%d.%s" % (trial, kind))
        p.write_bytes(data)
        args = ["-g", str(int(rng.integers(0, 500)))] + opts[trial % len(opts)]
        rc_r, out_r, _ = run(REF, args, str(p))
        rc_m, out_m, _ = run(MINE, args, str(p))
        assert rc_r == rc_m == 0
        assert out_m == out_r, (trial, args)


@pytest.mark.parametrize("golden", sorted(SPLIT))
def test_splitfa_matches_reference_output_byte_for_byte(golden):
    meta = SPLIT[golden]
    want = gzip.open(os.path.join(G, golden), "rb").read()
    rc, got, err = run(MINE_SPLIT, [os.path.join(G, meta["input"])] + meta["args"])
    assert rc == 0, err
    assert got == want and got.count(b">") == meta["records"]


def test_splitfa_pieces_are_what_psmc_split_uses(tmp_path):
    """the standalone tool and `psmc --split` (in memory, host/bootstrap.c) cut at the same places"""
    from psmc_b200 import psmcfa
    src = os.path.join(G, "cons_fq.s1.psmcfa.gz")
    rc, got, _ = run(MINE_SPLIT, [src, "2000"])
    p = tmp_path / "s.psmcfa"
    p.write_bytes(got)
    _, pieces = psmcfa.read_psmcfa(str(p))
    _, whole = psmcfa.read_psmcfa(src)
    want = []
    for s in whole:          # splitfa.c:24-29 restated
        i = 0
        while i < len(s):
            if len(s) - i < 2000 * 3 // 2:
                want.append(len(s) - i); break
            want.append(min(2000, len(s) - i)); i += 2000
    assert [len(x) for x in pieces] == want
    assert np.array_equal(np.concatenate(pieces), np.concatenate(whole))


@pytest.mark.skipif(not os.path.exists(REF_SPLIT), reason="reference utility not built here (oracle/_ref/splitfa)")
def test_splitfa_random_inputs_against_the_reference_binary(tmp_path):
    rng = np.random.default_rng(11)
    for trial in range(30):
        kind = "fq" if trial % 3 == 0 else "fa"
        p = tmp_path / ("t%d.%s" % (trial, kind))
        p.write_bytes(_random_input(rng, kind))
        t = str(int(rng.integers(1, 3000)))
        rc_r, out_r, _ = run(REF_SPLIT, [str(p), t])
        rc_m, out_m, _ = run(MINE_SPLIT, [str(p), t])
        assert rc_r == rc_m == 0 and out_m == out_r, (trial, t)
