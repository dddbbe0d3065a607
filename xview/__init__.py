"""Drop-in import path: `from xview.models import get_model` resolves to the B200-native
implementation in `modular_semantic_segmentation_b200.models`."""
