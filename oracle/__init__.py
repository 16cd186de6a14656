"""TEST INFRASTRUCTURE: CPU oracle of the tailored-avsr hot path (see ref_path.py).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
