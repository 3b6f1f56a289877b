"""Numeric-aware comparison of .psmc text (test infrastructure)."""
import re

NUM = re.compile(r"^[-+]?(\d+\.?\d*([eE][-+]?\d+)?|inf|nan)$")


def parse(path_or_text):
    txt = path_or_text if "\n" in path_or_text else open(path_or_text).read()
    return [l.rstrip("\n") for l in txt.splitlines()]


def fields(line):
    tag = line[:2]
    rest = line[3:] if len(line) > 2 else ""
    return tag, re.split(r"[\t ,]+", rest.strip()) if rest else []


def compare_rounds(got, want, tol):
    """tol: dict tag -> (rel, abs).  Header/CC/MM text must be identical; IT lines (objective-call counts) and QD are format-checked only."""
    assert len(got) == len(want), "line count differs: %d vs %d" % (len(got), len(want))
    worst = {}
    for i, (g, w) in enumerate(zip(got, want)):
        tg, fg = fields(g); tw, fw = fields(w)
        assert tg == tw, "line %d: tag %s vs %s" % (i, tg, tw)
        if tg in ("CC", "//", "RD") or (tg == "MM" and "C_pi" not in g):
            assert g == w, "line %d differs:\n%s\n%s" % (i, g, w)
            continue
        if tg in ("IT", "QD"):
            assert len(fg) == len(fw)
            continue
        assert len(fg) == len(fw), "line %d: field count" % i
        rel, ab = tol.get(tg, tol["*"])
        for a, b in zip(fg, fw):
            if NUM.match(a.rstrip(":")) and NUM.match(b.rstrip(":")):
                x, y = float(a), float(b)
                if x == y or (x != x and y != y):
                    continue
                err = abs(x - y)
                lim = max(rel * abs(y), ab)
                worst[tg] = max(worst.get(tg, 0.0), err / max(abs(y), 1e-300))
                assert err <= lim, "line %d (%s): %r vs %r (tol rel %g abs %g)\n%s\n%s" % (i, tg, x, y, rel, ab, g, w)
            else:
                assert a == b, "line %d: %r vs %r" % (i, a, b)
    return worst


RS_COLS = ["k", "t_k", "lambda_k", "pi_k", "sum_A_kl", "A_kk"]


def deviations(a, b, tags=("LK", "TR", "MT", "RS", "QD", "RI")):
    """largest relative deviation per tag (RS split by column) between two .psmc texts with the same line structure"""
    worst = {}
    assert len(a) == len(b), "line count differs: %d vs %d" % (len(a), len(b))
    for la, lb in zip(a, b):
        ta, fa = fields(la); tb, fb = fields(lb)
        assert ta == tb
        if ta not in tags:
            continue
        for j, (x, y) in enumerate(zip(fa, fb)):
            try:
                x = float(x); y = float(y)
            except ValueError:
                continue
            key = ta if ta != "RS" else "RS.%s" % RS_COLS[j]
            # the text carries six decimals (%lf): a difference of one unit in the last printed place is not a deviation
            diff = max(0.0, abs(x - y) - 1.5e-6)
            d = diff / abs(y) if y != 0 else diff
            worst[key] = max(worst.get(key, 0.0), d)
    return worst
