"""Baselines timed beside the product in bench.py (never imported by omniswarm_b200)."""
