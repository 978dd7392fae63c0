from . import mask as maskUtils  # noqa
class COCOeval: pass
