// slicer warm-up simulation: exact state vs speculated state at every segment start
#include <stdint.h>
#include <string.h>
#include <math.h>
typedef struct { double c; int last; } st_t;
static inline void step(st_t *s, int sg, double thr, double sps, double lam) {
	s->c += 1.0;
	if (s->c >= thr) s->c -= sps;
	if (sg != s->last) s->c *= lam;
	s->last = sg;
}
// by-crossing advance over [a,b), approx; mode 0 = float, 1 = double
static void approx_run(st_t *s, const uint8_t *sg, long a, long b, double thr, double sps, double lam, int mode) {
	if (mode == 0) {
		float c = (float)s->c, fthr=(float)thr, fsps=(float)sps, flam=(float)lam; int last = s->last; long pos = a;
		for (long i = a; i < b; i++) {
			if (sg[i] != last) { float u = c + (float)(i + 1 - pos); c = u >= fthr ? u - fsps*(floorf((u-fthr)/fsps)+1.0f) : u; c *= flam; pos = i + 1; last = sg[i]; }
		}
		float u = c + (float)(b - pos); c = u >= fthr ? u - fsps*(floorf((u-fthr)/fsps)+1.0f) : u;
		s->c = c; s->last = last;
	} else {
		double c = s->c; int last = s->last; long pos = a;
		for (long i = a; i < b; i++) {
			if (sg[i] != last) { double u = c + (double)(i + 1 - pos); c = u >= thr ? u - sps*(floor((u-thr)/sps)+1.0) : u; c *= lam; pos = i + 1; last = sg[i]; }
		}
		double u = c + (double)(b - pos); c = u >= thr ? u - sps*(floor((u-thr)/sps)+1.0) : u;
		s->c = c; s->last = last;
	}
}
// returns number of failed hand-offs among segment starts p = L, 2L, ... ; warm W samples total, exact tail X; mode of far part
// ncross > 0: adaptive: far part starts where ncross crossings precede p (capped at W)
long sim(const uint8_t *sg, long n, long L, long W, long X, int mode, int ncross, double sps, double lam, long *n_seg_out, double *avg_warm)
{
	double thr = sps / 2.0 - 0.5;
	st_t ex = {0.0, 1};
	long fails = 0, nseg = 0; double wsum = 0;
	long p = 0;
	for (long k = 1; k * L < n; k++) {
		long q = k * L;
		for (; p < q; p++) step(&ex, sg[p], thr, sps, lam);
		long a = q - W; if (a < 0) a = 0;
		if (ncross > 0) {
			int cnt = 0; long i = q - 1;
			for (; i > a; i--) { if (sg[i] != sg[i-1]) { if (++cnt >= ncross) break; } }
			a = i & ~31L; if (a < 0) a = 0;
		}
		st_t s = {0.0, 1};
		long mid = q - X; if (mid < a) mid = a;
		mid &= ~31L; if (mid < a) mid = a;
		approx_run(&s, sg, a, mid, thr, sps, lam, mode);
		for (long i = mid; i < q; i++) step(&s, sg[i], thr, sps, lam);
		wsum += (double)(q - a);
		nseg++;
		if (memcmp(&s.c, &ex.c, 8) != 0 || s.last != ex.last) fails++;
	}
	*n_seg_out = nseg; *avg_warm = wsum / (nseg ? nseg : 1);
	return fails;
}
