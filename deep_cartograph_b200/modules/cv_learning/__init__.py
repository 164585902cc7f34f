from .cv_calculator import (CVCalculator, LinearCalculator, PCACalculator, TICACalculator,  # noqa: F401
                            HTICACalculator, cv_calculators_map)
