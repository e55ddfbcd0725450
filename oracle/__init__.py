"""Test infrastructure: CPU restatement of the reference's decode path (see lattice_oracle.py).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package never imports it.
"""
