/* psmcfa.c -- .psmcfa reader.  Accepts what the reference accepts (cli.c:103-138 over kseq.h:173-217):
 * FASTA or FASTQ text, plain or gzip ("-" = stdin); a record starts at '>' or '@'; the name is the
 * header up to the first white space; every graphic character of the body is one bin until the next
 * '>', '@' or '+'; after '+' the rest of the line and as many quality characters as there were bins are
 * skipped.  Bins are mapped like conv_table (cli.c:15-32): ACGT0 -> 0, KMRSWY1 -> 1 (either case),
 * anything else -> 2.  A streaming state machine over 64 KB gzread blocks (no per-record line buffer). */
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <zlib.h>
#include "psmc_host.h"

static signed char g_conv[256];
static void conv_init(void)
{
	static int done = 0;
	const char *hom = "ACGT", *het = "KMRSWY";
	int i;
	if (done) return;
	for (i = 0; i < 256; ++i) g_conv[i] = 2;
	for (i = 0; hom[i]; ++i) g_conv[(int)hom[i]] = g_conv[tolower(hom[i])] = 0;
	for (i = 0; het[i]; ++i) g_conv[(int)het[i]] = g_conv[tolower(het[i])] = 1;
	g_conv['0'] = 0; g_conv['1'] = 1;
	done = 1;
}

typedef struct {
	gzFile fp;
	unsigned char buf[65536];
	int n, pos, eof;
} rd_t;

static int rd_getc(rd_t *r)
{
	if (r->pos >= r->n) {
		if (r->eof) return -1;
		r->n = gzread(r->fp, r->buf, sizeof(r->buf));
		r->pos = 0;
		if (r->n <= 0) { r->eof = 1; r->n = 0; return -1; }
	}
	return r->buf[r->pos++];
}

static void seq_push(psmch_seq_t *s, int64_t *cap, int c)
{
	signed char v;
	if (s->L + 1 >= *cap) {
		*cap = *cap ? *cap * 2 : 1 << 16;
		s->seq = (signed char*)realloc(s->seq, (size_t)*cap);
	}
	v = g_conv[c & 0xff];
	s->seq[s->L++] = v;
	if (v < 2) { ++s->L_e; if (v == 1) ++s->n_e; }
}

int psmch_read_psmcfa(const char *fn, psmch_seqs_t *out)
{
	rd_t *r = (rd_t*)calloc(1, sizeof(rd_t));
	int c, last = 0, cap_seqs = 0, i;
	conv_init();
	memset(out, 0, sizeof(*out));
	r->fp = strcmp(fn, "-") ? gzopen(fn, "r") : gzdopen(0, "r");
	if (r->fp == 0) { free(r); return -1; }
	for (;;) {
		psmch_seq_t s;
		int64_t cap = 0;
		char name[1024];
		int nl = 0;
		if (last == 0) { /* jump to the next header */
			while ((c = rd_getc(r)) != -1 && c != '>' && c != '@');
			if (c == -1) break;
		}
		memset(&s, 0, sizeof(s));
		/* name = up to the first white space; the rest of the header line is a comment */
		while ((c = rd_getc(r)) != -1 && !isspace(c))
			if (nl < (int)sizeof(name) - 1) name[nl++] = (char)c;
		name[nl] = 0;
		if (c == -1 && nl == 0) break;
		if (c != -1 && c != '\n') while ((c = rd_getc(r)) != -1 && c != '\n');
		/* body */
		while ((c = rd_getc(r)) != -1 && c != '>' && c != '+' && c != '@')
			if (isgraph(c)) seq_push(&s, &cap, c);
		last = (c == '>' || c == '@') ? c : 0;
		if (c == '+') { /* FASTQ: skip the '+' line, then L quality characters */
			int64_t q = 0;
			while ((c = rd_getc(r)) != -1 && c != '\n');
			if (c == -1) { free(s.seq); break; }             /* truncated: the reference stops reading here */
			while ((c = rd_getc(r)) != -1 && q < s.L)
				if (c >= 33 && c <= 127) ++q;
			if (q != s.L) { free(s.seq); break; }
		}
		if (s.seq == 0) s.seq = (signed char*)calloc(1, 1);
		s.name = strdup(name);
		if (out->n_seqs == cap_seqs) {
			cap_seqs = cap_seqs ? cap_seqs * 2 : 256;
			out->seqs = (psmch_seq_t*)realloc(out->seqs, sizeof(psmch_seq_t) * cap_seqs);
		}
		out->seqs[out->n_seqs++] = s;
		if (c == -1 && last == 0) break;
	}
	gzclose(r->fp);
	free(r);
	for (i = 0; i < out->n_seqs; ++i) { /* cli.c:133-137 */
		out->sum_n += out->seqs[i].n_e;
		out->sum_L += out->seqs[i].L_e;
	}
	return 0;
}

void psmch_free_seqs(psmch_seqs_t *s)
{
	int i;
	for (i = 0; i < s->n_seqs; ++i) { free(s->seqs[i].seq); free(s->seqs[i].name); }
	free(s->seqs);
	memset(s, 0, sizeof(*s));
}
