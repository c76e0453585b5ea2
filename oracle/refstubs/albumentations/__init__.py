"""Import stub for the absent `albumentations` dependency (test infrastructure only).

`od_datasets.py:9` does `import albumentations as A`; the transform builders
(`od_datasets.py:420-512`) are never called by the oracle harness, so any attribute
access raises to make accidental use loud.
"""


def __getattr__(name):
    raise AttributeError(f"albumentations stub: {name} is not available here")
