"""dedalus.data_objects (B200 backend): see api.py for the public names."""
