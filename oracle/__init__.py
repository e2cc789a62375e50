"""CPU oracles for the thermal hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this package; the product (our_first_climate_model_b200) never does.
"""
