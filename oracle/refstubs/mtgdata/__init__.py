"""Import stub for the absent `mtgdata` dependency (test infrastructure only).

The reference imports `ScryfallDataset, ScryfallImageType` (encoder_datasets.py:55)
and `mtgdata.scryfall.ScryfallBulkType, ScryfallCardFace` (:60).  Only the names are
needed to import `Mutate` / `SyntheticBgFgMtgImages.make_*`; nothing here is executed
on the path we validate.
"""
import enum


class ScryfallImageType(str, enum.Enum):
    small = "small"
    normal = "normal"


class ScryfallDataset:  # never instantiated by the oracle harness
    def __init__(self, *a, **k):
        raise RuntimeError("mtgdata stub: no Scryfall data in this environment")
