class PeftAdapterMixin:
    pass


class FluxLoraLoaderMixin:
    pass


class FromSingleFileMixin:
    pass
