"""Transform / PFFT base classes (test infrastructure only)."""
import numpy as np


class Transform:
    def __init__(self, xfftn, transfer, pencil):
        assert len(xfftn) == len(transfer) + 1 and len(pencil) == 2
        self._xfftn = tuple(xfftn)
        self._transfer = tuple(transfer)
        self._pencil = tuple(pencil)

    @property
    def input_array(self):
        return self._xfftn[0].input_array

    @property
    def output_array(self):
        return self._xfftn[-1].output_array

    @property
    def input_pencil(self):
        return self._pencil[0]

    @property
    def output_pencil(self):
        return self._pencil[1]

    def __call__(self, input_array=None, output_array=None, **kw):
        if input_array is not None:
            self.input_array[...] = input_array
        for i in range(len(self._transfer)):
            self._xfftn[i](**kw)
            self._transfer[i](self._xfftn[i].output_array, self._xfftn[i+1].input_array)
        self._xfftn[-1](**kw)
        if output_array is not None:
            output_array[...] = self.output_array
            return output_array
        return self.output_array


class PFFT:
    """Only the bookkeeping methods shenfun's TensorProductSpace inherits."""

    def destroy(self):
        self.subcomm.destroy()
        for trans in self.transfer:
            trans.destroy()

    def shape(self, forward_output=True):
        if forward_output is not True:
            return self.forward.input_pencil.subshape
        return self.forward.output_array.shape

    def local_shape(self, forward_output=True):
        return self.shape(forward_output)

    def local_slice(self, forward_output=True):
        if forward_output is not True:
            ip = self.forward.input_pencil
        else:
            ip = self.backward.input_pencil
        return tuple(slice(start, start+shape) for start, shape in zip(ip.substart, ip.subshape))

    def global_shape(self, forward_output=False):
        if forward_output:
            return self._output_shape
        return self._input_shape

    def dimensions(self):
        return len(self.forward.input_array.shape)

    def dtype(self, forward_output=False):
        """mpi4py_fft.mpifft.PFFT.dtype: dtype of the arrays before / after the forward transform."""
        if forward_output:
            return self.forward.output_array.dtype
        return self.forward.input_array.dtype
