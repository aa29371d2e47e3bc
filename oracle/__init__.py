"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the CUP3D hot path: ``refbind`` drives the real reference
(compiled from /root/reference into oracle/_ref/), ``portbind`` drives the C
restatement (oracle/cup_oracle.c).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the
product (cup3d_b200/) must never do so.
"""
