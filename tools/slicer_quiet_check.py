"""The claim behind SlicerChain::quiet_words (csrc/slicer.cu): without zero crossings the slicer's clock (slicer.py:77-81:
+1.0 per sample, -samples_per_symbol at the threshold) becomes EXACTLY periodic within one period when samples per symbol
is a dyadic rational -- the first pass through the top binade rounds it onto that binade's grid, after which every
operation is exact.  Checked here in CPython floats (the reference's arithmetic) from 20000 random start states per rate,
including states straight after lock multiplies (full mantissas) and tiny ones."""
import random


def run(c, sps, n):
	thr = sps / 2.0 - 0.5
	out = []
	for _ in range(n):
		c += 1.0
		if c >= thr:
			c -= sps
		out.append(c)
	return out


def period_of(sps):
	for k in range(1, 65):
		if sps * k == int(sps * k):
			return int(sps * k)
	return None


def check(sps, trials=20000, seed=1):
	random.seed(seed)
	P = period_of(sps)
	worst, fails = 0, 0
	for _ in range(trials):
		c = random.uniform(-sps / 2 - 1, sps / 2) * random.choice([1.0, 0.77, 0.77 ** 3, 1e-3, 1e-9])
		tr = run(c, sps, 6 * P + 200)
		ok_from = next((n0 for n0 in range(0, 4 * P) if all(tr[n + P] == tr[n] for n in range(n0, len(tr) - P))), None)
		if ok_from is None:
			fails += 1
		else:
			worst = max(worst, ok_from)
	return P, fails, worst


if __name__ == "__main__":
	for sps in (40.0, 36.75, 5.0, 10.0, 160.0, 18.375, 32.0, 80 / 3.0):
		dyadic = sps * 65536.0 == int(sps * 65536.0)
		P, fails, worst = check(sps, trials=20000 if dyadic else 200)
		print(f"sps {sps:.6g} dyadic {dyadic}: period {P} samples, start states that never become periodic {fails}, "
			f"periodic from sample <= {worst}")
