"""Field / dataset schema: the contract the embedding layer is constructed from.

Mirrors the reference's ``deepfm/data/schema.py:7-59`` (same class names, field
names, defaults and derived properties) so a schema built for the reference can be
handed to this package unchanged and vice versa.  The drop-in modules never test
``isinstance`` on these types: they read ``feature_type.value`` ("sparse" / "dense" /
"sequence"), so the reference's own enum is accepted as well.
"""

from __future__ import annotations

import dataclasses
import enum
from typing import Dict, List


class FeatureType(enum.Enum):
    SPARSE = "sparse"
    DENSE = "dense"
    SEQUENCE = "sequence"


@dataclasses.dataclass
class FieldSchema:
    name: str
    feature_type: FeatureType
    vocabulary_size: int = 0
    embedding_dim: int = 8
    group: str = ""
    max_length: int = 1
    combiner: str = "mean"


@dataclasses.dataclass
class DatasetSchema:
    fields: Dict[str, FieldSchema] = dataclasses.field(default_factory=dict)
    label_field: str = "label"

    def _of(self, kind: FeatureType) -> List[FieldSchema]:
        return [f for f in self.fields.values() if kind_of(f) == kind.value]

    @property
    def sparse_fields(self) -> List[FieldSchema]:
        return self._of(FeatureType.SPARSE)

    @property
    def dense_fields(self) -> List[FieldSchema]:
        return self._of(FeatureType.DENSE)

    @property
    def sequence_fields(self) -> List[FieldSchema]:
        return self._of(FeatureType.SEQUENCE)

    @property
    def num_fields(self) -> int:
        return len(self.fields)

    @property
    def total_embedding_dim(self) -> int:
        return sum(f.embedding_dim for f in self.fields.values())


def kind_of(field) -> str:
    """'sparse' | 'dense' | 'sequence' for this package's or the reference's FieldSchema."""
    ft = field.feature_type
    return ft.value if hasattr(ft, "value") else str(ft)
