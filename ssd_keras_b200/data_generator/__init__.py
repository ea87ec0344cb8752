"""The hand-off between the reference's data pipeline and the hot path: batch assembly for the encoder,
box transforms on the device and the inverse-transform helper.  Dataset parsing and image augmentation stay out of scope."""
