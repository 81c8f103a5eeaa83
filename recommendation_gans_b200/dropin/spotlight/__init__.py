"""Drop-in `spotlight` package: the subset of Spotlight that the implicit-MF fit/evaluate path of
Stamatios-Korres/recommendation_Gans uses, backed by the mfb200 CUDA library."""
__version__ = 'v0.1.5+mfb200'
