from .clip_encoders import (CustomImageEncoder, CustomTextEncoder, CustomVisionTransformer,
                            ImageEncoder, TextEncoder)
from .prompts_models import ImagePrefixModel, TextPrefixModel, UPTModel

__all__ = ["CustomImageEncoder", "CustomTextEncoder", "CustomVisionTransformer", "ImageEncoder",
           "TextEncoder", "ImagePrefixModel", "TextPrefixModel", "UPTModel"]
