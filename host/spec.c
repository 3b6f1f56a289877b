/* spec.c -- a helper thread that evaluates the M-step objective at ONE speculative trial point while the caller
 * evaluates another.  Hooke-Jeeves probes every coordinate as "+step, else -step" (kmin.c:48-66); both trial points are
 * known in advance, so the -step point is evaluated concurrently and used only if +step fails: the search path, the
 * number of counted calls and every accepted value are exactly those of the sequential search (the objective is a pure
 * function of the point), the wall time of a probe that fails in both directions -- most of them near convergence --
 * halves.  The helper spins while a search is running (an evaluation takes ~1.5 us: no blocking primitive is fast
 * enough) and sleeps on a condition variable between searches. */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "psmc_host.h"

#if defined(__x86_64__) || defined(__i386__)
#define PSMCH_CPU_RELAX() __builtin_ia32_pause()
#elif defined(__aarch64__)
#define PSMCH_CPU_RELAX() __asm__ __volatile__("yield")
#else
#define PSMCH_CPU_RELAX() ((void)0)
#endif

struct psmch_spec {
	psmch_func_t f;
	void *data;
	int n;
	double *x;
	volatile double value;
	volatile unsigned long long submitted, done; /* written by the caller / by the helper */
	volatile int active, quit;
	pthread_t th;
	pthread_mutex_t mu;
	pthread_cond_t cv;
};

static void *helper(void *arg)
{
	psmch_spec_t *s = (psmch_spec_t*)arg;
	unsigned long long seen = 0;
	for (;;) {
		pthread_mutex_lock(&s->mu);
		while (!s->active && !s->quit) pthread_cond_wait(&s->cv, &s->mu);
		pthread_mutex_unlock(&s->mu);
		if (s->quit) break;
		while (s->active) { /* a search is running: spin */
			const unsigned long long sub = __atomic_load_n(&s->submitted, __ATOMIC_ACQUIRE);
			if (sub != seen) {
				s->value = s->f(s->n, s->x, s->data);
				seen = sub;
				__atomic_store_n(&s->done, sub, __ATOMIC_RELEASE);
			} else {
				PSMCH_CPU_RELAX();
			}
		}
	}
	return 0;
}

psmch_spec_t *psmch_spec_start(psmch_func_t f, int n, void *data)
{
	psmch_spec_t *s = (psmch_spec_t*)calloc(1, sizeof(*s));
	if (s == 0) return 0;
	s->f = f; s->n = n; s->data = data;
	s->x = (double*)calloc(n > 0 ? n : 1, sizeof(double));
	pthread_mutex_init(&s->mu, 0);
	pthread_cond_init(&s->cv, 0);
	if (s->x == 0 || pthread_create(&s->th, 0, helper, s) != 0) { free(s->x); free(s); return 0; }
	return s;
}

void psmch_spec_begin(psmch_spec_t *s) /* a search starts: wake the helper */
{
	pthread_mutex_lock(&s->mu);
	s->active = 1;
	pthread_cond_signal(&s->cv);
	pthread_mutex_unlock(&s->mu);
}

void psmch_spec_end(psmch_spec_t *s) /* the search is over (nothing may be in flight): the helper goes to sleep */
{
	pthread_mutex_lock(&s->mu);
	s->active = 0;
	pthread_mutex_unlock(&s->mu);
}

void psmch_spec_submit(psmch_spec_t *s, const double *x)
{
	memcpy(s->x, x, sizeof(double) * s->n);
	__atomic_store_n(&s->submitted, s->submitted + 1, __ATOMIC_RELEASE);
}

double psmch_spec_wait(psmch_spec_t *s)
{
	while (__atomic_load_n(&s->done, __ATOMIC_ACQUIRE) != s->submitted) __builtin_ia32_pause();
	return s->value;
}

void psmch_spec_stop(psmch_spec_t *s)
{
	if (s == 0) return;
	pthread_mutex_lock(&s->mu);
	s->quit = 1; s->active = 0;
	pthread_cond_signal(&s->cv);
	pthread_mutex_unlock(&s->mu);
	pthread_join(s->th, 0);
	pthread_mutex_destroy(&s->mu);
	pthread_cond_destroy(&s->cv);
	free(s->x); free(s);
}
