from typing import Any

from langchain_core.documents import Document
from langchain_core.retrievers import BaseRetriever
from langchain_core.runnables import Runnable


class ContextualCompressionRetriever(BaseRetriever):
    """`base_retriever: RetrieverLike` -- pydantic rejects anything that is not a Runnable (reference app/rag.py:96-99)."""
    base_compressor: Any
    base_retriever: Runnable

    def _get_relevant_documents(self, query: str, *, run_manager: Any = None) -> list[Document]:
        docs = self.base_retriever.invoke(query)
        return list(self.base_compressor.compress_documents(docs, query)) if docs else []

    async def _aget_relevant_documents(self, query: str, *, run_manager: Any = None) -> list[Document]:
        docs = await self.base_retriever.ainvoke(query)
        return list(await self.base_compressor.acompress_documents(docs, query)) if docs else []
