"""CPU restatement of the reference's demod_chain path (oracle.c + oracle.py).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product path."""
