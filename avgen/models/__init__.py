from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)  # avgen.models.audio_encoders etc. resolve to the reference checkout
