"""Comparator plumbing (not product code): locating / importing the UNMODIFIED reference for baselines and live parity tests."""
