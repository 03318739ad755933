// Host-side exerciser of utils/cuda_vectors.h: every operator the header defines (+=, -=, *=, +, -, *
// against a vector and against a scalar, float4 and double2), evaluated on the host through the __host__ __device__
// definitions and printed one result component per line; tests/test_cuda_vectors_cpu.py compares with numpy.
#include <cstdio>

#include "../../utils/cuda_vectors.h"

static void show(const char *tag, double2 v) { std::printf("%s %.17g %.17g\n", tag, v.x, v.y); }
static void show(const char *tag, float4 v) { std::printf("%s %.9g %.9g %.9g %.9g\n", tag, v.x, v.y, v.z, v.w); }

int main()
{
    const double2 a = make_double2(1.25, -3.5), b = make_double2(0.1, 7.0);
    const float4 f = make_float4(1.5f, -2.25f, 0.3f, 8.0f), g = make_float4(0.7f, 4.0f, -1.1f, 0.125f);
    const double s = 2.5;
    const float t = -0.75f;
    double2 x;
    float4 y;
    show("d+v", a + b);
    show("d-v", a - b);
    show("d*v", a * b);
    show("d+s", a + s);
    show("d-s", a - s);
    show("d*s", a * s);
    x = a; x += b; show("d+=v", x);
    x = a; x -= b; show("d-=v", x);
    x = a; x *= b; show("d*=v", x);
    x = a; x += s; show("d+=s", x);
    x = a; x -= s; show("d-=s", x);
    x = a; x *= s; show("d*=s", x);
    show("f+v", f + g);
    show("f-v", f - g);
    show("f*v", f * g);
    show("f+s", f + t);
    show("f-s", f - t);
    show("f*s", f * t);
    y = f; y += g; show("f+=v", y);
    y = f; y -= g; show("f-=v", y);
    y = f; y *= g; show("f*=v", y);
    y = f; y += t; show("f+=s", y);
    y = f; y -= t; show("f-=s", y);
    y = f; y *= t; show("f*=s", y);
    // chained use as in the reference kernels: sum += d * d (benchmark01.cc:32), x += y (benchmark02.cc:29)
    x = make_double2(0.0, 0.0); x += a * a; x += b * b; show("dsq", x);
    return 0;
}
