from abc import ABC
from typing import Any


class Runnable(ABC):
    """What pydantic's isinstance check behind `RetrieverLike = Runnable[str, list[Document]]` looks for."""

    def invoke(self, input: Any, config: Any = None, **kwargs: Any) -> Any:
        raise NotImplementedError

    async def ainvoke(self, input: Any, config: Any = None, **kwargs: Any) -> Any:
        raise NotImplementedError
