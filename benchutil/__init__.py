"""Neutral helpers for bench.py, smoke() and the tools: synthetic rollouts / seeded weights (synth) and the
wiring of the product classes the way the reference starters wire them (harness).  Nothing here imports
oracle/ (the CPU restatement stays test infrastructure) and the product package never imports this."""
