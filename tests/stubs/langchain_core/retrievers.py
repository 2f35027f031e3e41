from abc import abstractmethod
from typing import Any, Optional

from pydantic import BaseModel, ConfigDict

from .documents import Document
from .runnables import Runnable


class BaseRetriever(BaseModel, Runnable):
    """RunnableSerializable[str, list[Document]] of langchain-core: invoke / ainvoke wrap the two hooks."""
    model_config = ConfigDict(arbitrary_types_allowed=True)
    tags: Optional[list] = None

    @abstractmethod
    def _get_relevant_documents(self, query: str, *, run_manager: Any = None) -> list[Document]: ...

    async def _aget_relevant_documents(self, query: str, *, run_manager: Any = None) -> list[Document]:
        return self._get_relevant_documents(query, run_manager=run_manager)

    def invoke(self, input: str, config: Any = None, **kwargs: Any) -> list[Document]:
        return self._get_relevant_documents(input, run_manager=None, **kwargs)

    async def ainvoke(self, input: str, config: Any = None, **kwargs: Any) -> list[Document]:
        return await self._aget_relevant_documents(input, run_manager=None, **kwargs)
