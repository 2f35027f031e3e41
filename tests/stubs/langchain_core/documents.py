from typing import Any, Optional

from pydantic import BaseModel, Field


class Document(BaseModel):
    page_content: str
    metadata: dict = Field(default_factory=dict)
    id: Optional[str] = None

    def __init__(self, page_content: str, **kwargs: Any) -> None:      # positional page_content, like langchain
        super().__init__(page_content=page_content, **kwargs)
