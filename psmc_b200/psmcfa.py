""".psmcfa text I/O for tests and benchmarks (the product reader is host/psmcfa.c).

Grammar as consumed by the reference (cli.c:103-138, kseq.h:173-217): FASTA records, every graphic
character of the body is one bin, mapped through conv_table (cli.c:15-32):
A C G T 0 -> 0 (hom);  K M R S W Y 1 -> 1 (het);  everything else -> 2 (missing), case-insensitive.
fq2psmcfa emits 60-column lines of T / K / N (utils/fq2psmcfa.c:118-135)."""
import gzip

import numpy as np

_CONV = np.full(256, 2, dtype=np.int8)
for ch in "ACGT":
    _CONV[ord(ch)] = 0; _CONV[ord(ch.lower())] = 0
_CONV[ord("0")] = 0
for ch in "KMRSWY":
    _CONV[ord(ch)] = 1; _CONV[ord(ch.lower())] = 1
_CONV[ord("1")] = 1
_OUT = np.frombuffer(b"TKN", dtype=np.uint8)


def write_psmcfa(path, seqs, names=None, width=60):
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "wb") as fp:
        for i, s in enumerate(seqs):
            name = names[i] if names else "chr%d" % (i + 1)
            fp.write((">%s\n" % name).encode())
            txt = _OUT[np.asarray(s, dtype=np.int8)]
            L = len(txt)
            nfull = L // width
            if nfull:
                body = np.empty((nfull, width + 1), dtype=np.uint8)
                body[:, :width] = txt[: nfull * width].reshape(nfull, width)
                body[:, width] = 10
                fp.write(body.tobytes())
            if L % width:
                fp.write(txt[nfull * width:].tobytes() + b"\n")


def read_psmcfa(path):
    op = gzip.open if str(path).endswith(".gz") else open
    names, seqs, cur = [], [], []
    with op(path, "rb") as fp:
        for line in fp:
            if line.startswith(b">"):
                if names:
                    seqs.append(np.concatenate(cur) if cur else np.zeros(0, dtype=np.int8))
                names.append(line[1:].split()[0].decode())
                cur = []
            else:
                b = np.frombuffer(line.strip(), dtype=np.uint8)
                b = b[(b > 32) & (b < 127)]
                cur.append(_CONV[b])
    if names:
        seqs.append(np.concatenate(cur) if cur else np.zeros(0, dtype=np.int8))
    return names, seqs
