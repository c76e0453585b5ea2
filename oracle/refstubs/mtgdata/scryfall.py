"""Stub of `mtgdata.scryfall` (names only) - see mtgdata/__init__.py."""
import enum


class ScryfallBulkType(str, enum.Enum):
    default_cards = "default_cards"


class ScryfallCardFace:
    pass
