from .time_step import TimeStepBase, RKBase, RK2mid, RK2trap, RK4, CrankNicholsonVisc
