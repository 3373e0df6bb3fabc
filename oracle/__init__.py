"""CPU oracle — TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package."""
