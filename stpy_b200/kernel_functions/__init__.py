"""Free-function twins of the Gram builders (mirror of stpy/kernel_functions/): same call signatures,
keyword protocol and return shapes, evaluated by the fused device Gram kernel."""
