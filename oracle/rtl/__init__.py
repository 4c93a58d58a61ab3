"""oracle/rtl — the reference's arithmetic one level below oracle/intfft_oracle.c: a behavioural DSP48E1/E2
primitive (dsp48.py), the reference's wrappers wired onto it port map by port map (netlist.py), and the two
butterfly entities on top of those (fly.py).  TEST INFRASTRUCTURE ONLY — never imported by the product."""
