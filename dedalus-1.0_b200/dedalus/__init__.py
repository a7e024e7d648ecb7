"""dedalus (B200-native): drop-in for the pseudospectral RHS + timestep hot path of
jsoishi/dedalus-1.0.  Same Python API (physics classes, StateData / FourierRepresentation,
time_stepping integrators); everything underneath is hand-written sm_100a CUDA reached through
the C ABI in include/ddl.h.  torch is used only as allocator and stream carrier."""
__version__ = "1.0+b200.r1"
