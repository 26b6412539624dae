"""`vsrd.configuration` — API of vsrd/configuration/configurator.py:7-164 that scripts/main.py:38 uses.

Configs are hierarchical: `Configurator.load(path/to/leaf/config.json)` merges every `config.json` found in
the parent directories (outermost first) with the leaf file; the same key may appear at two levels only with
the same value (configurator.py:116-164).  `gather` / `scatter` (factoring common keys up / pushing them back
down a config tree, configurator.py:10-113) are offline authoring tools and are not provided."""
import json
import os


class Configurator:

    @staticmethod
    def merge(*configs):
        """Deep union of dicts; a key present on both sides must hold equal leaves (AssertionError otherwise)."""
        def union(left, right):
            if not (isinstance(left, dict) and isinstance(right, dict)):
                assert left == right, f"conflicting config values: {left!r} vs {right!r}"
                return left
            merged = {key: (union(value, right[key]) if key in right else value) for key, value in left.items()}
            merged.update({key: value for key, value in right.items() if key not in left})
            return merged

        result = {}
        for config in configs:
            result = union(result, config)
        return result

    @staticmethod
    def load(filename):
        assert os.path.exists(filename), filename
        chain = []
        current = filename
        while os.path.exists(current):              # leaf first, then config.json of each ancestor directory
            with open(current) as file:
                chain.append(json.load(file))
            parent = os.path.dirname(os.path.dirname(current))
            following = os.path.join(parent, "config.json")
            if os.path.abspath(following) == os.path.abspath(current):
                break                                # reached the filesystem root
            current = following
        return Configurator.merge(*reversed(chain))
