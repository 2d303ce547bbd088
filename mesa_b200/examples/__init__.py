"""Device versions of the four reference example models named by BASELINE.json
(mesa/examples/basic/{schelling, conways_game_of_life, boltzmann_wealth_model, boid_flockers}); the advanced models
(wolf_sheep, epstein_civil_violence) are imported from their own modules."""
from .boid_flockers import BoidFlockers, BoidsScenario
from .boltzmann_wealth import BoltzmannScenario, BoltzmannWealth
from .conways_game_of_life import ConwaysGameOfLife
from .schelling import Schelling, SchellingScenario

__all__ = ["Schelling", "SchellingScenario", "ConwaysGameOfLife", "BoltzmannWealth", "BoltzmannScenario",
           "BoidFlockers", "BoidsScenario"]
