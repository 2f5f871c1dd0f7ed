"""Empty stand-in for imageio (video export helper of the reference's util.py; not on the hot path)."""
