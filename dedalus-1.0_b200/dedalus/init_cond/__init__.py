"""dedalus.init_cond (B200 backend): see api.py for the public names."""
