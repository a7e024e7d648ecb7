"""dedalus.utils (B200 backend): parallelism, timers, logging."""
