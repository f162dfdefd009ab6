"""Stand-in for the pieces of ConfigSpace used by bayesopt/generate_test_graphs.py:81-91,180-186."""
import random


class CategoricalHyperparameter:
    def __init__(self, name, choices):
        self.name = name
        self.choices = list(choices)


class ConfigurationSpace:
    def __init__(self):
        self._hp = []

    def add_hyperparameter(self, hp):
        self._hp.append(hp)
        return hp

    def sample_configuration(self):
        return {hp.name: random.choice(hp.choices) for hp in self._hp}
