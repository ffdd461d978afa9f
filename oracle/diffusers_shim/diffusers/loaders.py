class UNet2DConditionLoadersMixin:
    pass


class TextualInversionLoaderMixin:
    pass
