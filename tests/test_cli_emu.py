"""The drop-in `psmc` binary end to end on the CPU: host/psmc is linked against libpsmc_b200.so by name, so pointing the
loader at tests/emu's build of the same C ABI (the CUDA source executed by the SIMT emulation) runs the whole program --
CLI, .psmcfa reader, GPU-side orchestration, kernels' source, M-step, printer -- without a GPU, against the golden files
written by the unmodified reference.  Test infrastructure only (the product has no CPU path); `-m gpu` repeats it for real."""
import os
import shutil
import subprocess

import pytest

from psmc_text import compare_rounds, fields, parse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
PSMC = os.path.join(ROOT, "host", "psmc")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
TOL = {"*": (1e-6, 1e-6), "LK": (1e-7, 1e-6), "TR": (5e-5, 2e-6), "MT": (5e-5, 2e-6), "MM": (5e-5, 2e-6), "RS": (1e-3, 3e-6),
       "PA": (1e-3, 3e-6), "RI": (1e-3, 2e-7)}


@pytest.fixture(scope="module")
def emu_env(tmp_path_factory):
    if not os.path.exists(PSMC):
        pytest.skip("host/psmc not built")
    subprocess.run(["make", "-C", EMU_DIR], check=True, capture_output=True)
    d = tmp_path_factory.mktemp("emulib")
    shutil.copy(os.path.join(EMU_DIR, "libpsmc_b200_emu.so"), str(d / "libpsmc_b200.so"))
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = str(d) + os.pathsep + env.get("LD_LIBRARY_PATH", "")   # searched before the binary's RUNPATH
    env["PSMC_EMU_SMS"] = "2"
    return env


def run(args, env, out):
    r = subprocess.run([PSMC] + args + ["-o", out], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    return parse(out)


def test_emulated_cli_config1_matches_reference(emu_env, tmp_path):
    got = run(["-N5", "-t5", "-r1", "-p", "4+5*3+4", os.path.join(G, "c1.psmcfa.gz")], emu_env, str(tmp_path / "c1.psmc"))
    compare_rounds(got, parse(os.path.join(G, "c1.psmc")), TOL)


def test_emulated_cli_64_states_exact_qd(emu_env, tmp_path):
    """64 states, ragged contigs, --exact-qd: every line including QD against the reference's output"""
    args = ["-N4", "-t15", "-r5", "-p", "4+25*2+4+6", "--exact-qd", os.path.join(G, "small64.psmcfa.gz")]
    got = run(args, emu_env, str(tmp_path / "s64.psmc"))
    want = parse(os.path.join(G, "small64.psmc"))
    compare_rounds(got, want, TOL)
    qd_g = [[float(x) for x in fields(l)[1] if x != "->"] for l in got if l.startswith("QD")]
    qd_w = [[float(x) for x in fields(l)[1] if x != "->"] for l in want if l.startswith("QD")]
    assert len(qd_g) == len(qd_w) > 1
    for a, b in zip(qd_g, qd_w):
        for x, y in zip(a, b):
            assert abs(x - y) <= 2e-5 * abs(y) + 2e-6, (a, b)


def test_emulated_replicates_with_exact_qd(emu_env, tmp_path):
    """--replicates together with --exact-qd: the shared bootstrap context must spill the backward rows too"""
    args = ["-N2", "-t15", "-r5", "-p", "4+25*2+4+6", "--exact-qd", "--split=300", "--replicates", "2", "--seed", "5",
            os.path.join(G, "small64.psmcfa.gz")]
    got = run(args, emu_env, str(tmp_path / "rep.psmc"))
    assert sum(1 for l in got if l.startswith("QD")) == 6          # 2 replicates x (round 0 + 2 iterations)
    assert sum(1 for l in got if l.startswith("RD")) == 6


def test_emulated_batched_replicates_equal_one_at_a_time(emu_env, tmp_path):
    """--replicates with the batched E-step (all replicates share every launch, the default) prints what the
    one-replicate-at-a-time scheme (--batch 1) prints; with a fixed chunk length the E-step statistics are the same bits,
    so the text is identical"""
    base = ["-N3", "-t15", "-r5", "-p", "4+25*2+4+6", "--split=300", "--replicates", "3", "--seed", "11", "--chunk", "200",
            os.path.join(G, "small64.psmcfa.gz")]
    one = run(base + ["--batch", "1"], emu_env, str(tmp_path / "one.psmc"))
    bat = run(base, emu_env, str(tmp_path / "bat.psmc"))
    two = run(base + ["--batch", "2"], emu_env, str(tmp_path / "two.psmc"))     # a ragged last batch
    assert sum(1 for l in bat if l.startswith("RD")) == 12
    assert bat == one
    assert two == one


def test_emulated_cli_decode_modes_match_reference(emu_env, tmp_path):
    """-d (runs compacted on the device), -D (float rows, hand-rolled %.4f on a thread pool) and -s against the
    reference's own output for the same fixed parameters"""
    import gzip
    import numpy as np
    fa, par = os.path.join(G, "c1.psmcfa.gz"), os.path.join(G, "c1_params.txt")
    got = run(["-N0", "-i", par, "-d", "--chunk", "700", fa], emu_env, str(tmp_path / "d.psmc"))
    want = parse(os.path.join(G, "c1_decode.psmc"))
    assert [l for l in got if l.startswith("TC")] == [l for l in want if l.startswith("TC")]
    dc_g = [l for l in got if l.startswith("DC")]; dc_w = [l for l in want if l.startswith("DC")]
    same = sum(a == b for a, b in zip(dc_g, dc_w))
    assert len(dc_g) == len(dc_w) and same >= 0.995 * len(dc_w), (len(dc_g), len(dc_w), same)
    assert [l for l in got if l[:2] not in ("TC", "DC")] == [l for l in want if l[:2] not in ("TC", "DC")]
    got = run(["-N0", "-i", par, "-D", "--chunk", "700", fa], emu_env, str(tmp_path / "D.psmc"))
    want = gzip.open(os.path.join(G, "c1_fulldecode.psmc.gz"), "rt").read().splitlines()
    df_g = [l for l in got if l.startswith("DF")]; df_w = [l for l in want if l.startswith("DF")]
    assert len(df_g) == len(df_w) == 10000
    A = np.array([[float(x) for x in l.split("\t")[1:]] for l in df_g])
    B = np.array([[float(x) for x in l.split("\t")[1:]] for l in df_w])
    assert np.array_equal(A[:, 0], B[:, 0])
    assert np.max(np.abs(A[:, 1] - B[:, 1])) <= 2e-6           # recombination probability, %lf
    assert np.max(np.abs(A[:, 2:] - B[:, 2:])) <= 1.0001e-4    # posteriors, %.4f
    assert np.mean(A[:, 2:] == B[:, 2:]) > 0.999               # (float rows: the last printed digit may differ at a rounding boundary)
    assert all(len(l.split("\t")[2]) >= 8 for l in df_g[:50])  # %lf keeps six decimals
