from .volume_average import VolumeAverageSet, volume_average
