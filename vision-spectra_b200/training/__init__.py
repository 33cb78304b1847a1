"""Trainer-side caller of the hot path (reference training/base.py:379-416)."""
