"""CPU oracle for intfftk_b200 — TEST INFRASTRUCTURE ONLY (see oracle/intfft_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  PARITY UNPINNED by the reference (no upstream golden vectors)."""
