"""imported by sagan/dataset.py:5; only its tfds-backed loader (not on the pinned path) uses it."""
