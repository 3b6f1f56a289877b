/* splitfa -- cut the records of a .psmcfa into pieces of `trunk` bins for bootstrapping (README:49-62 of the reference).
 *
 * Same command line and the same output, byte for byte, as the reference's utils/splitfa.c (tests/test_fq2psmcfa.py): a record
 * is cut every `trunk` characters, except that a remainder shorter than 1.5 trunks stays in one piece (splitfa.c:24-29); pieces
 * are named <record>_<k>, k from 1, and printed in 60 columns (splitfa.c:8-18).  `host/psmc --split[=T]` applies the same rule
 * in memory (host/bootstrap.c, psmch_split); this program exists for workflows that keep the split file.
 */
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct { gzFile fp; unsigned char buf[1 << 16]; int n, pos, eof; } rd_t;

static inline int rd_getc(rd_t *r)
{
	if (r->pos >= r->n) {
		if (r->eof) return -1;
		r->n = gzread(r->fp, r->buf, sizeof(r->buf));
		r->pos = 0;
		if (r->n <= 0) { r->eof = 1; r->n = 0; return -1; }
	}
	return r->buf[r->pos++];
}

static void put_piece(const char *name, const char *s, int64_t beg, int64_t end, int id, char **buf, size_t *cap)
{
	const size_t need = strlen(name) + 32 + (size_t)(end - beg) + (size_t)(end - beg) / 60 + 2;
	char *p;
	int64_t i;
	if (need > *cap) { *cap = need * 2; *buf = (char*)realloc(*buf, *cap); }
	p = *buf + sprintf(*buf, ">%s_%d\n", name, id);
	for (i = beg; i < end; ++i) {
		if (i > beg && (i - beg) % 60 == 0) *p++ = '\n';
		*p++ = s[i];
	}
	*p++ = '\n';
	fwrite(*buf, 1, p - *buf, stdout);
}

int main(int argc, char *argv[])
{
	int trunk = 500000, last = 0, c;
	rd_t *r;
	char *seq = 0, *name = 0, *out = 0;
	int64_t cap = 0;
	size_t ocap = 0, ncap = 0;
	if (argc < 2) {
		fprintf(stderr, "Usage: splitfa <in.fa> [trunk_size=%d]\n", trunk);
		return 1;
	}
	if (argc >= 3) trunk = atoi(argv[2]);
	if (trunk <= 0) { fprintf(stderr, "[E::main] trunk_size must be positive\n"); return 1; }
	r = (rd_t*)calloc(1, sizeof(rd_t));
	r->fp = strcmp(argv[1], "-") ? gzopen(argv[1], "r") : gzdopen(0, "r");
	if (r->fp == 0) { fprintf(stderr, "[E::main] cannot open '%s'\n", argv[1]); return 1; }
	for (;;) { /* record grammar of kseq.h:172-217 */
		int64_t len = 0, i, q;
		size_t nl = 0;
		int k = 0, truncated = 0;
		if (last == 0) {
			while ((c = rd_getc(r)) != -1 && c != '>' && c != '@');
			if (c == -1) break;
		}
		while ((c = rd_getc(r)) != -1 && !isspace(c)) {
			if (nl + 2 > ncap) name = (char*)realloc(name, ncap = ncap ? ncap * 2 : 64);
			name[nl++] = (char)c;
		}
		if (c == -1 && nl == 0) break;
		if (name == 0) name = (char*)calloc(ncap = 64, 1);
		name[nl] = 0;
		if (c != -1 && c != '\n') while ((c = rd_getc(r)) != -1 && c != '\n');
		while ((c = rd_getc(r)) != -1 && c != '>' && c != '+' && c != '@') {
			if (!isgraph(c)) continue;
			if (len == cap) seq = (char*)realloc(seq, cap = cap ? cap * 2 : 1 << 16);
			seq[len++] = (char)c;
		}
		last = (c == '>' || c == '@') ? c : 0;
		if (c == '+') { /* FASTQ: the qualities are read and dropped; a short quality string ends the input */
			while ((c = rd_getc(r)) != -1 && c != '\n');
			if (c == -1) break;
			q = 0;
			while ((c = rd_getc(r)) != -1 && q < len)
				if (c >= 33 && c <= 127) ++q;
			last = 0;
			if (q != len) truncated = 1;
		}
		if (truncated) break;
		for (i = 0; i < len; i += trunk) {
			if (len - i < (int64_t)trunk * 3 / 2) { /* the remainder stays whole */
				put_piece(name, seq, i, len, ++k, &out, &ocap);
				break;
			}
			put_piece(name, seq, i, i + trunk < len ? i + trunk : len, ++k, &out, &ocap);
		}
	}
	gzclose(r->fp);
	free(r); free(seq); free(name); free(out);
	return 0;
}
