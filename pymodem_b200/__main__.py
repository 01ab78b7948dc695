"""python -m pymodem_b200 <config json> <sound file>

The reference's command line (pymodem.py:25-183) with the per-chain process fan-out (pymodem.py:140-166) replaced by
one GPU engine call: JSON-lines config -> chain objects (chain_builder) -> libpymodem_b200.so -> PacketMeta lists in
config order -> PacketMetaArray.CalcCRCs / Correlate(address_distance = sample_rate / 40) -> report.  Exit codes follow
the reference: 2 usage, 3 config file, 4 audio file.  The WAV (16-bit PCM mono, as the reference's sample files) is
memory-mapped and copied once, into page-locked memory (engine.read_wav_pinned); the engine DMAs it to the GPU in chunks
that overlap the front-end kernels.  There is no CPU fallback."""
import json
import sys
import time



def main(argv):
	if len(argv) != 3:
		print("Not enough arguments. Usage: python3 -m pymodem_b200 <config json> <sound file>")
		return 2
	try:
		with open(argv[1], 'r') as f:
			stack_plan = [json.loads(line) for line in f if line.strip()]
	except (OSError, ValueError):
		print('Unable to open config json file.')
		return 3
	try:
		from .engine import read_wav_pinned
		input_sample_rate, input_audio = read_wav_pinned(argv[2])      # the one copy of the samples: file -> pinned memory
	except Exception:
		print('Unable to open audio file.')
		return 4
	from .modems_codecs import chain_builder, chain_execute
	from .modems_codecs.packet_meta import PacketMetaArray, ReportStyle
	print("Building processing stacks from config json")
	demod_stack, report_stack = [], []
	for number, line in enumerate(stack_plan, 1):
		kind = line.get('object_type')
		if kind == 'demod_chain':
			try:
				demod_stack.append(chain_builder.build_chain(input_sample_rate, line))
				print(f"Line {number}: {line['object_name']}")
			except (KeyError, NotImplementedError, ValueError) as exc:
				print(f"Skipping chain in line {number}: {exc}")
		elif kind == 'report':
			report_stack.append((line.get('object_name', 'report'), ReportStyle(line.get('options', {}))))
	if not demod_stack:
		print("No usable demod_chain in the config.")
		return 3
	print("Executing demod stack plan.")
	start_time = time.time()
	decoded_datas = chain_execute.process_chains(demod_stack, input_audio)
	print("Correlating results.")
	results = PacketMetaArray()
	for decoded_data in decoded_datas:
		results.add(decoded_data)
	results.CalcCRCs()
	results.Correlate(address_distance=input_sample_rate / 40)
	for name, style in report_stack or [("report", None)]:
		print(f"Generating {name}")
		print(results.Report(style))
	print(f"Elapsed time: {round(time.time() - start_time, 2)} seconds.")
	return 0


if __name__ == "__main__":
	sys.exit(main(sys.argv))
