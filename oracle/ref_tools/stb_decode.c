/*
 * TEST INFRASTRUCTURE (oracle/_ref): runs the reference's own image decoder — stb_image.h as vendored under
 * /root/reference/src/lib/external/stb, compiled where it lies (nothing is copied into this repo) — so that the
 * from-scratch PNG / JPEG / Radiance decoders of vviewer_b200/host can be checked texel for texel.
 *
 *   stb_decode <file> [flip]      -> stdout: "W H C\n" + W*H*4 bytes RGBA (C = channels in the file), rows as stbi returns
 *   stb_decode --hdr <file> [flip] -> stdout: "W H C\n" + W*H*4 floats
 * Mirrors the reference call sites: stbi_load(..., STBI_rgb_alpha) (AssimpLoadModel.cpp:190, 211-223, core/Image.cpp:9-43)
 * with stbi_set_flip_vertically_on_load(true) when `flip` is given (core/Image.cpp:39).
 */
#define STB_IMAGE_IMPLEMENTATION
#include <stb_image.h>
#include <stdio.h>
#include <string.h>

int main(int argc, char **argv) {
    int hdr = 0, a = 1;
    if (argc > 1 && !strcmp(argv[1], "--hdr")) {
        hdr = 1;
        a = 2;
    }
    if (argc <= a) {
        fprintf(stderr, "usage: stb_decode [--hdr] file [flip]\n");
        return 2;
    }
    if (argc > a + 1 && !strcmp(argv[a + 1], "flip")) stbi_set_flip_vertically_on_load(1);
    int w, h, c;
    if (hdr) {
        float *d = stbi_loadf(argv[a], &w, &h, &c, STBI_rgb_alpha);
        if (!d) return 1;
        printf("%d %d %d\n", w, h, c);
        fwrite(d, sizeof(float), (size_t)w * h * 4, stdout);
    } else {
        unsigned char *d = stbi_load(argv[a], &w, &h, &c, STBI_rgb_alpha);
        if (!d) return 1;
        printf("%d %d %d\n", w, h, c);
        fwrite(d, 1, (size_t)w * h * 4, stdout);
    }
    return 0;
}
