"""tf.dtypes subset."""
bool = "bool"
float32 = "float32"
