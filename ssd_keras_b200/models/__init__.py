"""Model builders (``ssd_300`` / ``ssd_512`` / ``build_model``) and ``load_model`` for the files ``SSDModel.save`` writes."""


def load_model(filepath, custom_objects=None, compile=True):
    """``keras.models.load_model(filepath, custom_objects={...})`` for files written by :meth:`SSDModel.save`: rebuilds the model
    with the builder call stored in the file's ``model_config`` attribute and loads the weights.  ``custom_objects`` (the reference
    passes ``AnchorBoxes`` / ``L2Normalization`` / ``DecodeDetections`` / ``compute_loss``, ``ssd300_inference.ipynb``) is accepted
    and ignored -- those layers are built in.  A full-model file saved by Keras itself describes an arbitrary Keras graph, which
    this package does not interpret: build the model with ``ssd_300(...)`` and ``load_weights(filepath, by_name=True)`` instead
    (the weights of such files are read)."""
    import json
    from ..misc_utils.hdf5_lite import read_attributes
    from . import keras_ssd300, keras_ssd512, keras_ssd7
    raw = read_attributes(str(filepath)).get('model_config')
    if raw is None:
        raise ValueError('%s has no model_config attribute: it is a weights file; build the model and call load_weights()' % filepath)
    if hasattr(raw, 'tobytes'):
        raw = raw.tobytes()
    if isinstance(raw, bytes):
        raw = raw.rstrip(b'\x00').decode()
    cfg = json.loads(raw)
    if cfg.get('class_name') != 'SSDModel':
        raise ValueError('%s was saved by Keras (class_name %r): rebuild the architecture with ssd_300 / ssd_512 / build_model and '
                         'use load_weights(path, by_name=True)' % (filepath, cfg.get('class_name')))
    builders = {'ssd_300': keras_ssd300.ssd_300, 'ssd_512': keras_ssd512.ssd_512, 'build_model': keras_ssd7.build_model}
    b = cfg['config']['builder']
    if b not in builders:
        raise ValueError('unknown builder %r in %s' % (b, filepath))
    model = builders[b](**cfg['config']['kwargs'])
    model.load_weights(str(filepath), by_name=True)
    return model
