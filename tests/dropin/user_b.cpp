// Second translation unit including the same headers (a header-only library must link from many).
#include "avir_b200.h"
#include "lancir_b200.h"

#include <cstdint>

int resize_with_other_unit(const uint8_t* in, int w, int h, uint8_t* out, int nw, int nh) {
    avir::CLancIR L; // upstream lancir.h usage: ImageResizer.resizeImage( InBuf, 640, 480, OutBuf, 1024, 768, 3 )
    return L.resizeImage(in, w, h, out, nw, nh, 3);
}
