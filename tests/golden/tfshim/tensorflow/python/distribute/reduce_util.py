"""imported by sagan/layers.py:3, never used there."""
