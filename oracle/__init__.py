"""TEST INFRASTRUCTURE ONLY. CPU restatements of the reference algorithms on the hot path.
Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
as the checker — never by the product package u2seg_b200/."""
