"""dedalus.physics (B200 backend): see api.py for the public names."""
