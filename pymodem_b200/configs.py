"""Config builders for the workloads BASELINE.json names.  The reference keeps
these as JSON-lines files under configs/ (one object per line, pymodem.py:36-40);
here they are generated so that tests and bench.py do not need the reference
checkout.  The produced objects are exactly what json.loads() yields for the
reference's files (values are strings, as its StringOptionsRetune expects)."""


def _chain(name, modem, slicer, stream, codec):
	return {"object_name": name, "object_type": "demod_chain", "modem": modem, "slicer": slicer,
		"stream": stream, "codec": codec}


def afsk_1200_ax25_super_opt():
	"""The many-chain AFSK 1200 fan-out (reference configs/afsk_1200_ax25_super_opt.json:1-9):
	one 1600/1800 Hz chain with span-1.0 correlators, seven 1300/2100 Hz chains with
	span-1.5 correlators and space_gain 1.25 ... 2.75, all NRZI + AX.25, lock_rate 0.77."""
	slicer = {"type": "binary", "config": "1200", "options": {"lock_rate": "0.77"}}
	stream = {"type": "lfsr", "options": {"poly": "0x3", "invert": "True"}}
	lines = [_chain("AFSK 1200 AX.25 1600/1800 sg 1.0",
		{"type": "afsk", "config": "1200", "options": {"space_gain": "1.0", "mark_freq": "1600.0", "space_freq": "1800.0"}},
		slicer, stream, {"type": "ax25"})]
	for sg in ("1.25", "1.5", "1.75", "2.0", "2.25", "2.5", "2.75"):
		lines.append(_chain(f"AFSK 1200 AX.25 1300/2100 sg {sg}",
			{"type": "afsk", "config": "1200", "options": {"space_gain": sg, "mark_freq": "1300.0",
				"space_freq": "2100.0", "correlator_span": "1.5"}},
			slicer, stream, {"type": "ax25"}))
	lines.append({"object_name": "Decoded header report", "object_type": "report",
		"options": {"style": "decoded_headers", "destination": "std_out"}})
	return lines


def demod_chains(lines):
	return [l for l in lines if l.get("object_type") == "demod_chain"]
