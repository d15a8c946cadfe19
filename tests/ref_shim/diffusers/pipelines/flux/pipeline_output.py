from dataclasses import dataclass
from typing import Any

from ...utils import BaseOutput


@dataclass
class FluxPipelineOutput(BaseOutput):
    images: Any = None
