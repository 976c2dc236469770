from .PyHashGrid import PyHashGrid
from .PyHashGridBG import PyHashGridBG
