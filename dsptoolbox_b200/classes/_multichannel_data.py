"""Channel bookkeeping shared by data containers (reference classes/_multichannel_data.py:6-118,
only the members the accelerated path touches)."""

from __future__ import annotations

from abc import ABC, abstractmethod


class MultichannelData(ABC):
    @abstractmethod
    def _get_data(self):
        ...

    @property
    def number_of_channels(self) -> int:
        return self._get_data().shape[-1]

    def __len__(self):
        return self._get_data().shape[0]
