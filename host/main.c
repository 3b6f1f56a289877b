/* main.c -- the drop-in `psmc` driver: same round structure as main.c:6-34 of the reference
 * (RD 0 + parameters, then N times: EM iteration, RD i, parameters; then optional decoding). */
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "psmc_host.h"

static double rnd48(void) { return drand48(); }

int main(int argc, char *argv[])
{
	psmch_opts_t o;
	psmch_seqs_t sq;
	psmch_em_t em;
	psmch_space_t hdr;
	int i, rc;
	if (psmch_parse_cli(argc, argv, &o) != 0) return 1;
	srand48(o.seed >= 0 ? o.seed : (long)(time(0) ^ getpid())); /* main.c:11 */
	if (o.n_replicates == 0) psmch_print_header(&o, 0, 0, 0);
	if (o.pre_fn && psmch_read_param(&o, 0) != 0) return 1;
	if (psmch_space_init(&hdr, o.pattern ? o.pattern : "4+5*3+4", 0, o.alpha0) != 0) {
		fprintf(stderr, "psmc: bad pattern '%s'\n", o.pattern);
		return 1;
	}
	if (o.n_replicates == 0) psmch_print_header(&o, &hdr, 0, 1);
	psmch_space_free(&hdr);
	if (psmch_read_psmcfa(o.in_fn, &sq) != 0) {
		fprintf(stderr, "psmc: cannot read '%s'\n", o.in_fn);
		return 1;
	}
	if (o.split_len > 0 && psmch_split(&sq, o.split_len) != 0) return 1;
	if (o.n_replicates > 0) { /* README:57-62 in one process: every replicate prints its own complete .psmc text */
		rc = psmch_bootstrap_run(&o, &sq, o.n_replicates, o.slots);
		psmch_free_seqs(&sq);
		if (o.fpout != stdout) fclose(o.fpout);
		return rc != 0;
	}
	if (o.is_bootstrap) psmch_resample(&sq, rnd48);
	psmch_print_header(&o, 0, &sq, 2);
	if (psmch_em_init(&em, &o, &sq, rnd48) != 0) return 1;
	fprintf(o.fpout, "RD\t0\n");
	psmch_print_round(&o, &em, &sq, o.fpout);
	for (i = 0; i < o.n_iters; ++i) {
		if ((rc = psmch_em_iterate(&em, o.fpout)) != 0) return 1;
		if (o.verbose) fprintf(stderr, "[psmc-b200] iter %d: E-step %.2f ms, M-step %.2f ms (%d objective calls)\n", i + 1, em.t_estep_ms, em.t_mstep_ms, em.hj_calls);
		fprintf(o.fpout, "RD\t%d\n", i + 1);
		psmch_print_round(&o, &em, &sq, o.fpout);
	}
	if ((o.flag & PSMCH_F_DECODE) || (o.flag & PSMCH_F_PROB))
		if (psmch_decode(&o, &em, &sq, o.fpout) != 0) return 1;
	psmch_em_free(&em);
	psmch_free_seqs(&sq);
	if (o.fpout != stdout) fclose(o.fpout);
	return 0;
}
