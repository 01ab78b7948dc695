import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import Golden, as_tuples
from pymodem_b200.modems_codecs import chain_builder
from pymodem_b200.sharded import run_linked_local
g = Golden("afsk1200_superopt_48k")
stack = [chain_builder.build_chain(g.sample_rate, l) for l in g.lines if l.get("object_type") == "demod_chain"]
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
got, info = run_linked_local(stack, g.audio(), world, tail_bits=2048, segment_len=4096, warmup_len=16384)
print(info['verified'], as_tuples(got) == g.all_packets())
