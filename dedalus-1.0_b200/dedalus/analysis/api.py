from .volume_average import VolumeAverageSet, volume_average
from .analysis_set import AnalysisSet, AnalysisTask, Snapshot, TrackMode, PowerSpectrum, VolumeAverage
