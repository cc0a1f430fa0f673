from .heston import HestonModel, HestonParameters, OptionGreeks, PricingResult  # noqa: F401
