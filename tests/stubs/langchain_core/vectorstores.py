import asyncio
from abc import ABC, abstractmethod
from typing import Any, Iterable, Optional

from pydantic import ConfigDict, Field

from .documents import Document
from .retrievers import BaseRetriever


class VectorStore(ABC):
    @abstractmethod
    def add_texts(self, texts: Iterable[str], metadatas: Optional[list] = None, *, ids: Optional[list] = None,
                  **kwargs: Any) -> list: ...

    @property
    def embeddings(self):
        return None

    def delete(self, ids: Optional[list] = None, **kwargs: Any) -> Optional[bool]:
        raise NotImplementedError

    async def adelete(self, ids: Optional[list] = None, **kwargs: Any) -> Optional[bool]:
        return await asyncio.get_running_loop().run_in_executor(None, lambda: self.delete(ids, **kwargs))

    def add_documents(self, documents: list, **kwargs: Any) -> list:
        if "ids" not in kwargs:
            ids = [d.id for d in documents]
            if any(ids):
                kwargs["ids"] = ids
        return self.add_texts([d.page_content for d in documents], [d.metadata for d in documents], **kwargs)

    async def aadd_documents(self, documents: list, **kwargs: Any) -> list:
        return await asyncio.get_running_loop().run_in_executor(None, lambda: self.add_documents(documents, **kwargs))

    @abstractmethod
    def similarity_search(self, query: str, k: int = 4, **kwargs: Any) -> list: ...

    async def asimilarity_search(self, query: str, k: int = 4, **kwargs: Any) -> list:
        return await asyncio.get_running_loop().run_in_executor(None, lambda: self.similarity_search(query, k=k, **kwargs))

    @classmethod
    @abstractmethod
    def from_texts(cls, texts: list, embedding: Any, metadatas: Optional[list] = None, **kwargs: Any): ...

    def _get_retriever_tags(self) -> list:
        return [self.__class__.__name__]

    def as_retriever(self, **kwargs: Any) -> "VectorStoreRetriever":
        tags = (kwargs.pop("tags", None) or []) + self._get_retriever_tags()
        return VectorStoreRetriever(vectorstore=self, tags=tags, **kwargs)


class VectorStoreRetriever(BaseRetriever):
    model_config = ConfigDict(arbitrary_types_allowed=True)
    vectorstore: VectorStore                     # pydantic: isinstance(value, VectorStore)
    search_type: str = "similarity"
    search_kwargs: dict = Field(default_factory=dict)

    def _get_relevant_documents(self, query: str, *, run_manager: Any = None, **kwargs: Any) -> list[Document]:
        assert self.search_type == "similarity"
        return self.vectorstore.similarity_search(query, **{**self.search_kwargs, **kwargs})

    async def _aget_relevant_documents(self, query: str, *, run_manager: Any = None, **kwargs: Any) -> list[Document]:
        assert self.search_type == "similarity"
        return await self.vectorstore.asimilarity_search(query, **{**self.search_kwargs, **kwargs})
