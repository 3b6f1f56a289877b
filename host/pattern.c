/* pattern.c -- "-p" pattern parser.  Behaviour of cli.c:66-99 (psmc_parse_pattern):
 * terms separated by '+', each "len" or "rep*len"; a group is a run of `len` consecutive intervals
 * sharing one lambda.  n = (sum of all lengths) - 1, n_free = number of groups. */
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include "psmc_host.h"

int psmch_parse_pattern(const char *pattern, int *n_free, int **par_map)
{
	const char *p;
	int cap = 16, ng = 0, *glen, total = 0, i, j, pos, *map;
	long total_guard = 0;
	if (pattern == 0) return -1;
	for (p = pattern; *p; ++p)
		if (!isdigit((unsigned char)*p) && *p != '*' && *p != '+') return -1; /* the reference asserts (cli.c:74) */
	glen = (int*)malloc(sizeof(int) * cap);
	p = pattern;
	for (;;) {
		long a = strtol(p, (char**)&p, 10), rep = 1, len = a;
		if (*p == '*') { rep = a; ++p; len = strtol(p, (char**)&p, 10); }
		if (rep < 0 || ng + rep > 255) { free(glen); return -1; }            /* stack depth assert (cli.c:81) */
		if (len < 0 || len > PSMCH_MAX_INTERVALS || total_guard + rep * len > PSMCH_MAX_INTERVALS) { free(glen); return -1; } /* no int overflow, sane cap */
		total_guard += rep * len;
		if (ng + rep > cap) { while (ng + rep > cap) cap <<= 1; glen = (int*)realloc(glen, sizeof(int) * cap); }
		for (i = 0; i < rep; ++i) glen[ng++] = (int)len;
		if (*p == '+') { ++p; continue; }
		break;
	}
	for (i = 0; i < ng; ++i) total += glen[i];
	if (total < 1) { free(glen); return -1; }
	map = (int*)malloc(sizeof(int) * total);
	for (i = 0, pos = 0; i < ng; ++i)
		for (j = 0; j < glen[i]; ++j) map[pos++] = i;
	free(glen);
	if (n_free) *n_free = ng;
	if (par_map) *par_map = map; else free(map);
	return total - 1;
}

int psmch_space_init(psmch_space_t *sp, const char *pattern, int diverg, double alpha0)
{
	memset(sp, 0, sizeof(*sp));
	sp->n = psmch_parse_pattern(pattern, &sp->n_free, &sp->par_map);
	if (sp->n < 0) return -1;
	sp->pattern = strdup(pattern);
	sp->diverg = diverg;
	sp->alpha0 = alpha0;
	sp->n_params = sp->n_free + PSMCH_N_PARAMS + (diverg ? 1 : 0); /* core.c:26 */
	return 0;
}

void psmch_space_free(psmch_space_t *sp)
{
	free(sp->par_map); free(sp->pattern); free(sp->inp_ti);
	memset(sp, 0, sizeof(*sp));
}
