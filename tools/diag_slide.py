import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from test_gpu_parity import Golden, engine, build_stack
for tag in ("afsk1200_superopt_48k", "afsk1200_ax25_44k1"):
	g = Golden(tag)
	res = {}
	for slide in (0, 1):
		eng = engine(build_stack(g.sample_rate, g.lines), slide_correlator=slide, keep_soft=1)
		eng.run_raw(g.audio())
		res[slide] = [eng.soft(ci).astype(np.float64) for ci in range(g.n_chains)]
		eng.close()
	for ci in range(g.n_chains):
		if f"c{ci}_soft_dec" not in g.z: continue
		rms = float(g.z[f"c{ci}_soft_rms"]); ref = g.z[f"c{ci}_soft_dec"]
		e0 = np.abs(res[0][ci][::97] - ref); e1 = np.abs(res[1][ci][::97] - ref)
		d = np.abs(res[0][ci] - res[1][ci])
		print(tag, ci, 'rms', rms, 'direct max/rms err', e0.max(), np.sqrt((e0**2).mean()), 'slide', e1.max(), np.sqrt((e1**2).mean()), 'argmax', e1.argmax()*97, 'ref there', ref[e1.argmax()], 'd max', d.max(), d.argmax(), 'absmax soft', np.abs(res[0][ci]).max())
		i = d.argmax(); u = (i // 16) * 16
		print('   around', res[0][ci][u:u+16] - res[1][ci][u:u+16])
