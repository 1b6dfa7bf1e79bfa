"""Parity checker for nka_b200 -- TEST INFRASTRUCTURE ONLY (see nka_oracle.c header)."""
