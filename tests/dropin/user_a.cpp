// A user translation unit written against upstream's documented API (README "Usage Information"),
// with only the include changed -- compiled by tests/test_dropin.py, never part of the library.
#include "avir_b200.h"
#include "lancir_b200.h"

#include <cstdint>
#include <cstdio>
#include <exception>
#include <vector>

int resize_with_other_unit(const uint8_t* in, int w, int h, uint8_t* out, int nw, int nh); // user_b.cpp

int main(int argc, char** argv) {
    const int W = 640, H = 480, NW = 1024, NH = 768, C = 3;
    std::vector<uint8_t> in((size_t)W * H * C), out((size_t)NW * NH * C), out2((size_t)NW * NH * C);
    uint32_t s = 12345;
    for (auto& v : in) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; v = (uint8_t)(s >> 24); }
    if (argc > 2) { // the test's input image
        FILE* f = std::fopen(argv[2], "rb");
        if (!f || std::fread(in.data(), 1, in.size(), f) != in.size()) return 6;
        std::fclose(f);
    }
    try {
        // upstream README: avir::CImageResizer<> ImageResizer( 8 ); ImageResizer.resizeImage( InBuf, 640, 480, 0, OutBuf, 1024, 768, 3, 0 );
        avir::CImageResizer<> ImageResizer(8);
        ImageResizer.resizeImage(in.data(), W, H, 0, out.data(), NW, NH, C, 0);
        // the documented variations: parameter presets, Vars, the SIMD classes, 16-bit and float buffers
        // (compiled and linked always; run when the program is started without arguments)
        if (argc <= 1) {
        avir::CImageResizerVars Vars;
        Vars.UseSRGBGamma = true;
        avir::CImageResizer<avir::fpclass_float4> R4(8, 0, avir::CImageResizerParamsUltra());
        R4.resizeImage(in.data(), W, H, 0, out2.data(), NW, NH, C, 0, &Vars);
        std::vector<float> fin((size_t)W * H * C, 0.5f), fout((size_t)NW * NH * C);
        avir::CImageResizer<avir::fpclass_float8_dil> R8(16);
        R8.resizeImage(fin.data(), W, H, 0, fout.data(), NW, NH, C, 0);
        std::vector<uint16_t> win((size_t)W * H * C, 1000), wout((size_t)NW * NH * C);
        avir::CImageResizer<> R16(16);
        R16.resizeImage(win.data(), W, H, 0, wout.data(), NW, NH, C, 0);
        }
        if (resize_with_other_unit(in.data(), W, H, out2.data(), NW, NH) != NH) return 4;
    } catch (const std::exception& e) {
        std::printf("threw: %s\n", e.what()); // (no CUDA device: the library has no CPU fallback)
        return 3;
    }
    if (argc > 1) {
        FILE* f = std::fopen(argv[1], "wb");
        if (!f || std::fwrite(out.data(), 1, out.size(), f) != out.size()) return 5;
        std::fclose(f);
    }
    std::printf("ok\n");
    return 0;
}
