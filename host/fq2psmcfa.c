/* fq2psmcfa -- diploid consensus FASTQ/FASTA -> .psmcfa, the producer of the E-step's input.
 *
 * Same options and the same output, byte for byte, as the reference's utils/fq2psmcfa.c (tests/test_fq2psmcfa.py holds it
 * to the unmodified utility's output), written from its behaviour:
 *   record grammar        kseq.h:172-217 (FASTA or FASTQ, multi-line, name = header up to the first white space, a record with
 *                         a truncated quality string ends the input)
 *   base classes          utils/fq2psmcfa.c:13-32 (IUPAC -> 4-bit set; lower case = masked), here built from the IUPAC
 *                         definitions instead of a literal table
 *   masks                 -q (fq2psmcfa.c:69-72), -x (64-67), -v (73-77), -n (78-88), -c (89-101), -C (102-111)
 *   binning               fq2psmcfa.c:115-127: one character per -s bases: N if more than 90 % of a FULL block is masked
 *                         (also for the short last block), else K if any base is a two-allele code, else T
 *   record filter, layout fq2psmcfa.c:129-136: >= 20 % and >= -g unmasked bases; 60 characters per line
 *
 * Structure: per record a class byte per base (bit set + mask flag), one pass per selected mask rule, one binning pass;
 * records are independent, so -p N formats N records at a time on N threads while the main thread reads (the reference is
 * single-threaded; its cost at 3 Gbp is gzip and this loop).  Output order is the input order.
 */
#include <ctype.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#define MASKED 0x10 /* flag on a class byte: treated as missing (the reference lower-cases the base) */
#define MISSING 15

typedef struct {
	int min_qual, min_good, block, mask_par, rule; /* rule: 0 none, 'v', 'n', 'c', 'C' */
	int threads;
} opt_t;

typedef struct {
	char *name;
	unsigned char *cls; /* per base: 4-bit allele set | MASKED */
	char *qual;         /* or NULL (FASTA) */
	int64_t len, cap, qcap;
	char *out;          /* formatted record or NULL (dropped) */
	size_t out_len;
} rec_t;

static unsigned char allele_set[256]; /* A=1 C=2 G=4 T=8; IUPAC unions; X=0; everything else 15 */
static unsigned char is_lower_tab[256];

static void tables_init(void)
{
	static const struct { char c; unsigned char set; } iupac[] = {
		{'A', 1}, {'C', 2}, {'G', 4}, {'T', 8}, {'M', 1 | 2}, {'R', 1 | 4}, {'W', 1 | 8}, {'S', 2 | 4}, {'Y', 2 | 8}, {'K', 4 | 8},
		{'V', 1 | 2 | 4}, {'H', 1 | 2 | 8}, {'D', 1 | 4 | 8}, {'B', 2 | 4 | 8}, {'N', 15}, {'X', 0}};
	int i;
	memset(allele_set, MISSING, sizeof(allele_set));
	for (i = 0; i < (int)(sizeof(iupac) / sizeof(iupac[0])); ++i) {
		allele_set[(unsigned char)iupac[i].c] = iupac[i].set;
		allele_set[(unsigned char)tolower(iupac[i].c)] = iupac[i].set;
	}
	for (i = 0; i < 256; ++i) is_lower_tab[i] = (i >= 'a' && i <= 'z');
}

/* ---- buffered reader --------------------------------------------------------------------------- */
typedef struct {
	gzFile fp;
	unsigned char buf[1 << 16];
	int n, pos, eof;
} rd_t;

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

/* one record; returns 1, or 0 at the end of the input (also after a record whose quality string is short, like the reference) */
static int read_record(rd_t *r, int *last, rec_t *s, int mask_lower_input)
{
	int c, nl = 0, ncap = 64;
	(void)mask_lower_input;
	if (*last == 0) {
		while ((c = rd_getc(r)) != -1 && c != '>' && c != '@');
		if (c == -1) return 0;
	}
	s->name = (char*)malloc(ncap);
	while ((c = rd_getc(r)) != -1 && !isspace(c)) {
		if (nl + 2 > ncap) s->name = (char*)realloc(s->name, ncap *= 2);
		s->name[nl++] = (char)c;
	}
	s->name[nl] = 0;
	if (c == -1 && nl == 0) { free(s->name); s->name = 0; return 0; }
	if (c != -1 && c != '\n') while ((c = rd_getc(r)) != -1 && c != '\n');
	s->len = 0;
	while ((c = rd_getc(r)) != -1 && c != '>' && c != '+' && c != '@') {
		if (!isgraph(c)) continue;
		if (s->len == s->cap) {
			s->cap = s->cap ? s->cap * 2 : 1 << 16;
			s->cls = (unsigned char*)realloc(s->cls, s->cap);
		}
		s->cls[s->len++] = (unsigned char)(allele_set[c] | (is_lower_tab[c] ? MASKED : 0));
	}
	*last = (c == '>' || c == '@') ? c : 0;
	s->qual = 0;
	if (c == '+') {
		int64_t q = 0;
		while ((c = rd_getc(r)) != -1 && c != '\n');
		if (c == -1) return -1;
		if (s->qcap < s->len + 1) s->qcap = s->len + 1;
		s->qual = (char*)malloc(s->qcap);
		while ((c = rd_getc(r)) != -1 && q < s->len)
			if (c >= 33 && c <= 127) s->qual[q++] = (char)c;
		*last = 0;
		if (q != s->len) return -1;
	}
	return 1;
}

/* ---- masks and binning (one record) -------------------------------------------------------------- */
static inline int is_tv_het(int c) { return c == (1 | 2) || c == (1 | 8) || c == (2 | 4) || c == (4 | 8); } /* M W S K */

static void format_record(rec_t *s, const opt_t *o)
{
	const int64_t len = s->len;
	unsigned char *cl = s->cls;
	int64_t i, n_good = 0, n_bins, l = 0;
	char *bins;
	if (o->mask_par && (strcmp(s->name, "X") == 0 || strcmp(s->name, "chrX") == 0)) { /* pseudo-autosomal regions, b36 coordinates */
		static const int64_t par[2][2] = {{1, 2709520}, {154584237, 154913754}};
		int k;
		for (k = 0; k < 2; ++k)
			for (i = par[k][0] - 1; i < par[k][1] && i < len; ++i) cl[i] |= MASKED;
	}
	if (s->qual)
		for (i = 0; i < len; ++i)
			if (s->qual[i] - 33 < o->min_qual) cl[i] |= MASKED;
	if (o->rule) { /* the rules look at the CALLED code of a base and of its predecessor, whether or not those are already masked */
		int pre = -1;
		for (i = 0; i < len; ++i) {
			const int c = cl[i] & 15;
			const int cpg = i > 0 && (c == 4 || c == (1 | 4)) && (pre == 2 || pre == (2 | 8)); /* G or R after C or Y */
			switch (o->rule) {
			case 'v': if (c == (1 | 4) || c == (2 | 8)) cl[i] |= MASKED; break; /* transitions R, Y */
			case 'n':
				if (cpg) { cl[i] |= MASKED; cl[i - 1] |= MASKED; }
				else if (is_tv_het(c)) cl[i] |= MASKED;
				break;
			case 'c':
				if (is_tv_het(c)) cl[i] |= MASKED;
				else if (i > 0 && pre == (2 | 8) && c != 4 && c != (1 | 4)) cl[i - 1] |= MASKED; /* a Y not followed by G/R */
				else if (i > 0 && c == (1 | 4) && pre != 2 && pre != (2 | 8)) cl[i] |= MASKED;    /* an R not preceded by C/Y */
				break;
			case 'C': if (cpg) { cl[i] |= MASKED; cl[i - 1] |= MASKED; } break;
			}
			pre = c;
		}
	}
	n_bins = (len + o->block - 1) / o->block + 1;
	bins = (char*)malloc(n_bins + 1);
	{
		int64_t b0;
		for (b0 = 0; b0 < len || b0 == 0; b0 += o->block) { /* (an empty record still yields one bin, as in the reference) */
			const int64_t b1 = b0 + o->block < len ? b0 + o->block : len;
			int n_miss = 0, het = 0;
			for (i = b0; i < b1; ++i) {
				const int c = (cl[i] & MASKED) ? MISSING : (cl[i] & 15);
				if (c == MISSING) ++n_miss;
				else if (c == 3 || c == 5 || c == 6 || c == 9 || c == 10 || c == 12) het = 1; /* exactly two alleles */
			}
			n_good += (b1 - b0) - n_miss;
			bins[l++] = ((float)n_miss / o->block > 0.9) ? 'N' : (het ? 'K' : 'T'); /* (float, against the double 0.9: 90 of 100 is NOT above it) */
		}
	}
	s->out = 0;
	s->out_len = 0;
	if (len > 0 && (double)n_good / len >= 0.2 && n_good >= o->min_good) {
		const size_t nlen = strlen(s->name);
		char *p = s->out = (char*)malloc(nlen + l + l / 60 + 8);
		*p++ = '>';
		memcpy(p, s->name, nlen); p += nlen;
		for (i = 0; i < l; ++i) {
			if (i % 60 == 0) *p++ = '\n';
			*p++ = bins[i];
		}
		*p++ = '\n';
		s->out_len = p - s->out;
	}
	free(bins);
}

static void rec_release(rec_t *s)
{
	free(s->name); free(s->cls); free(s->qual); free(s->out);
	memset(s, 0, sizeof(*s));
}

typedef struct { rec_t *s; const opt_t *o; } job_t;
static void *job_run(void *a) { job_t *j = (job_t*)a; format_record(j->s, j->o); return 0; }

static int usage(const opt_t *o)
{
	fprintf(stderr, "Usage: fq2psmcfa [-cnvx] [-q %d] [-g %d] [-s %d] [-p threads] <in.fq>\n", o->min_qual, o->min_good, o->block);
	return 1;
}

int main(int argc, char *argv[])
{
	opt_t o = {10, 10000, 100, 0, 0, 1};
	int c, n_rules = 0, last = 0, done = 0;
	rd_t *r;
	tables_init();
	while ((c = getopt(argc, argv, "q:xg:s:vncCp:")) >= 0) {
		switch (c) {
		case 'q': o.min_qual = atoi(optarg); break;
		case 'x': o.mask_par = 1; break;
		case 'g': o.min_good = atoi(optarg); break;
		case 's': o.block = atoi(optarg); break;
		case 'p': o.threads = atoi(optarg); break;
		case 'v': case 'n': case 'c': case 'C':
			if (o.rule != c) ++n_rules;
			o.rule = c;
			break;
		default: return usage(&o);
		}
	}
	if (n_rules > 1) {
		fprintf(stderr, "[E::main] only one of the options -c, -n, -v and -C can be applied\n");
		return 2;
	}
	if (argc == optind) return usage(&o);
	if (o.block <= 0) { fprintf(stderr, "[E::main] -s must be positive\n"); return 2; }
	if (o.threads < 1) o.threads = 1;
	if (o.threads > 64) o.threads = 64;
	r = (rd_t*)calloc(1, sizeof(rd_t));
	r->fp = strcmp(argv[optind], "-") ? gzopen(argv[optind], "r") : gzdopen(0, "r");
	if (r->fp == 0) { fprintf(stderr, "[E::main] cannot open '%s'\n", argv[optind]); free(r); return 1; }
	gzbuffer(r->fp, 1 << 20);
	while (!done) { /* waves of up to `threads` records: read on this thread, format concurrently, write in order */
		rec_t wave[64];
		job_t jobs[64];
		pthread_t th[64];
		int n = 0, k;
		memset(wave, 0, sizeof(wave));
		while (n < o.threads) {
			const int rc = read_record(r, &last, &wave[n], 0);
			if (rc <= 0) { rec_release(&wave[n]); done = 1; break; }
			jobs[n].s = &wave[n]; jobs[n].o = &o;
			if (o.threads > 1) pthread_create(&th[n], 0, job_run, &jobs[n]);
			else job_run(&jobs[n]);
			++n;
		}
		for (k = 0; k < n; ++k) {
			if (o.threads > 1) pthread_join(th[k], 0);
			if (wave[k].out) fwrite(wave[k].out, 1, wave[k].out_len, stdout);
			rec_release(&wave[k]);
		}
		fflush(stdout);
	}
	gzclose(r->fp);
	free(r);
	return 0;
}
