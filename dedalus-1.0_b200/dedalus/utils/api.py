from .function_count import counts
from .timer import Timer, timer
from .parallelism import com_sys, swap_indices, load_all, reduce_mean, reduce_sum, reduce_min, reduce_max
from .restart import restart, load_all_data
