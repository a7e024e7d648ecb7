"""Run-time configuration, mirroring dedalus/config.py:29-56 of the reference: a module-level
ConfigParser ``decfg`` with the same sections, keys and defaults, read from
``~/.dedalus/config`` and ``./dedalus.cfg``.

Differences: ``FFT.method`` selects nothing here -- 'fftw', 'numpy' and 'cuda' all mean the
one CUDA backend (there is no CPU path); anything else raises NotImplementedError when a
representation is built, as in representations.py:309-310.
"""
import configparser
import os

decfg = configparser.ConfigParser()

_DEFAULTS = {
    "FFT": {"method": "cuda", "dealiasing": "2/3 cython"},
    # slab decomposition over the ranks of the process group (one per GPU).  ky_layout:
    # 'block' = the reference's contiguous ky slabs (representations.py:231-233);
    # 'cyclic' = rank r owns ky rows r, r+P, ... (balanced under 2/3 dealiasing).
    # exchange: 'peer' (stores into the peers' memory fused into the passes), 'p2p' (copy
    # engines), 'collective' (torch.distributed all_to_all_single)
    "parallel": {"ky_layout": "block", "exchange": "peer"},
    "physics": {"use_tracer": "False", "boussinesq_direction": "z"},
    "forcing": {},
    "utils": {"loglevel": "warning", "loadplugins": "False", "pluginfilename": "dedalus_plugins.py"},
    "analysis": {"snapshot_space": "xspace", "snapshot_axis": "z", "snapshot_index": "middle",
                 "snapshot_units": "True", "snapshot_dpi": "100", "snapshot_cmap": "Spectral_r",
                 "powerspectrum_dpi": "100"},
}
for _sec, _kv in _DEFAULTS.items():
    decfg.add_section(_sec)
    for _k, _v in _kv.items():
        decfg.set(_sec, _k, _v)

decfg.read([os.path.expanduser("~/.dedalus/config"), "dedalus.cfg"])
