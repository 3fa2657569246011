"""Mirror of the reference package A2SB/audio_transforms (see transforms.py)."""
