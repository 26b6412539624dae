from .sinusoidal_encoder import SinusoidalEncoder  # noqa: F401
