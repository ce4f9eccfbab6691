"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference algorithms used as the parity checker.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product path (mmfn_b200/) never does.
"""
