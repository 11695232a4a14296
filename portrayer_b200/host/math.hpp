// Host-side math with the conventions of the reference's src/math.rs (type
// aliases over vek 0.9.8, f64 everywhere).  vek's sources are not available
// offline; the operation orders chosen here are documented in DESIGN.md
// ("parity unpinned at ULP level" items) and are the SAME ones the oracle and
// the CUDA kernels use: separate multiply and add, left-to-right sums.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>

namespace portrayer {

constexpr double EPSILON = 0.00001;  // src/math.rs:15
constexpr double GAMMA = 2.2;        // src/math.rs:20
constexpr double INFINITY_F64 = std::numeric_limits<double>::infinity();
constexpr double PI = 3.14159265358979323846264338327950288;

struct Vec3 {
    double x = 0, y = 0, z = 0;
    constexpr Vec3() = default;
    constexpr Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    // vek: a scalar broadcasts to all components (Vec3::from(f64), `.scaled(100.0)`)
    constexpr Vec3(double s) : x(s), y(s), z(s) {}

    static constexpr Vec3 zero() { return {0, 0, 0}; }
    static constexpr Vec3 up() { return {0, 1, 0}; }
    static constexpr Vec3 down() { return {0, -1, 0}; }
    static constexpr Vec3 right() { return {1, 0, 0}; }
    static constexpr Vec3 forward_rh() { return {0, 0, -1}; }
    static constexpr Vec3 back_rh() { return {0, 0, 1}; }
    static constexpr Vec3 unit_x() { return {1, 0, 0}; }
    static constexpr Vec3 unit_y() { return {0, 1, 0}; }
    static constexpr Vec3 unit_z() { return {0, 0, 1}; }

    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }

    Vec3 operator+(Vec3 o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vec3 operator-(Vec3 o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vec3 operator*(Vec3 o) const { return {x * o.x, y * o.y, z * o.z}; }
    Vec3 operator/(Vec3 o) const { return {x / o.x, y / o.y, z / o.z}; }
    Vec3 operator*(double s) const { return {x * s, y * s, z * s}; }
    Vec3 operator/(double s) const { return {x / s, y / s, z / s}; }
    Vec3 operator-() const { return {-x, -y, -z}; }
    Vec3& operator+=(Vec3 o) { *this = *this + o; return *this; }
    bool operator==(Vec3 o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(Vec3 o) const { return !(*this == o); }

    double sum() const { return x + y + z; }
    double dot(Vec3 o) const { return x * o.x + y * o.y + z * o.z; }
    double magnitude_squared() const { return dot(*this); }
    double magnitude() const { return std::sqrt(magnitude_squared()); }
    Vec3 normalized() const { return *this / magnitude(); }
    Vec3 cross(Vec3 b) const { return {y * b.z - z * b.y, z * b.x - x * b.z, x * b.y - y * b.x}; }

    static Vec3 partial_min(Vec3 a, Vec3 b) { return {pmin(a.x, b.x), pmin(a.y, b.y), pmin(a.z, b.z)}; }
    static Vec3 partial_max(Vec3 a, Vec3 b) { return {pmax(a.x, b.x), pmax(a.y, b.y), pmax(a.z, b.z)}; }

  private:
    // vek::ops::partial_min: `if a <= b { a } else { b }`
    static double pmin(double a, double b) { return a <= b ? a : b; }
    static double pmax(double a, double b) { return a >= b ? a : b; }
};
inline Vec3 operator*(double s, Vec3 v) { return {s * v.x, s * v.y, s * v.z}; }

struct Rgb {
    double r = 0, g = 0, b = 0;
    constexpr Rgb() = default;
    constexpr Rgb(double r_, double g_, double b_) : r(r_), g(g_), b(b_) {}
    static constexpr Rgb black() { return {0, 0, 0}; }
    static constexpr Rgb white() { return {1, 1, 1}; }
    static constexpr Rgb red() { return {1, 0, 0}; }
    static constexpr Rgb green() { return {0, 1, 0}; }
    static constexpr Rgb blue() { return {0, 0, 1}; }
    Rgb operator+(Rgb o) const { return {r + o.r, g + o.g, b + o.b}; }
    Rgb operator*(Rgb o) const { return {r * o.r, g * o.g, b * o.b}; }
    Rgb operator*(double s) const { return {r * s, g * s, b * s}; }
    Rgb operator/(double s) const { return {r / s, g / s, b / s}; }
    bool operator==(Rgb o) const { return r == o.r && g == o.g && b == o.b; }
};
inline Rgb operator*(double s, Rgb c) { return {s * c.r, s * c.g, s * c.b}; }

struct Uv {
    double u = 0, v = 0;
};

// Radians newtype, src/math.rs:55-72
struct Radians {
    double value = 0;
    static Radians from_degrees(double deg) { return Radians{deg * (PI / 180.0)}; }  // f64::to_radians
    static Radians from_radians(double rad) { return Radians{rad}; }
    double get() const { return value; }
};

// 4x4 matrix; m[r][c] is the mathematical element (row r, column c) — the
// value vek's column-major Mat4 exposes as `cols[c][r]`.
struct Mat4 {
    double m[4][4];

    static Mat4 identity() {
        Mat4 r{};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = (i == j) ? 1.0 : 0.0;
        return r;
    }
    static Mat4 scaling_3d(Vec3 v) {
        Mat4 r = identity();
        r.m[0][0] = v.x; r.m[1][1] = v.y; r.m[2][2] = v.z;
        return r;
    }
    static Mat4 translation_3d(Vec3 v) {
        Mat4 r = identity();
        r.m[0][3] = v.x; r.m[1][3] = v.y; r.m[2][3] = v.z;
        return r;
    }
    static Mat4 rotation_x(double a) {
        double c = std::cos(a), s = std::sin(a);
        Mat4 r = identity();
        r.m[1][1] = c; r.m[1][2] = -s; r.m[2][1] = s; r.m[2][2] = c;
        return r;
    }
    static Mat4 rotation_y(double a) {
        double c = std::cos(a), s = std::sin(a);
        Mat4 r = identity();
        r.m[0][0] = c; r.m[0][2] = s; r.m[2][0] = -s; r.m[2][2] = c;
        return r;
    }
    static Mat4 rotation_z(double a) {
        double c = std::cos(a), s = std::sin(a);
        Mat4 r = identity();
        r.m[0][0] = c; r.m[0][1] = -s; r.m[1][0] = s; r.m[1][1] = c;
        return r;
    }
    // GL-convention right-handed look-at (world -> view)
    static Mat4 look_at_rh(Vec3 eye, Vec3 center, Vec3 up) {
        Vec3 f = (center - eye).normalized();
        Vec3 s = f.cross(up).normalized();
        Vec3 u = s.cross(f);
        Mat4 r = identity();
        r.m[0][0] = s.x; r.m[0][1] = s.y; r.m[0][2] = s.z; r.m[0][3] = -s.dot(eye);
        r.m[1][0] = u.x; r.m[1][1] = u.y; r.m[1][2] = u.z; r.m[1][3] = -u.dot(eye);
        r.m[2][0] = -f.x; r.m[2][1] = -f.y; r.m[2][2] = -f.z; r.m[2][3] = f.dot(eye);
        return r;
    }

    Mat4 operator*(const Mat4& o) const {
        Mat4 r{};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j] + m[i][3] * o.m[3][j];
        return r;
    }
    // vek builder methods compose on the LEFT: m.scaled_3d(v) == S * m.
    // Pinned by src/bounding_box.rs:185-195 (rotated_cube_bounds_60).
    Mat4 scaled_3d(Vec3 v) const { return scaling_3d(v) * *this; }
    Mat4 translated_3d(Vec3 v) const { return translation_3d(v) * *this; }
    Mat4 rotated_x(double a) const { return rotation_x(a) * *this; }
    Mat4 rotated_y(double a) const { return rotation_y(a) * *this; }
    Mat4 rotated_z(double a) const { return rotation_z(a) * *this; }

    Mat4 transposed() const {
        Mat4 r{};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = m[j][i];
        return r;
    }

    // General 4x4 inverse by Laplace expansion (2x2 sub-determinants).
    Mat4 inverted() const {
        const double m00 = m[0][0], m01 = m[0][1], m02 = m[0][2], m03 = m[0][3];
        const double m10 = m[1][0], m11 = m[1][1], m12 = m[1][2], m13 = m[1][3];
        const double m20 = m[2][0], m21 = m[2][1], m22 = m[2][2], m23 = m[2][3];
        const double m30 = m[3][0], m31 = m[3][1], m32 = m[3][2], m33 = m[3][3];
        const double s0 = m00 * m11 - m10 * m01, s1 = m00 * m12 - m10 * m02, s2 = m00 * m13 - m10 * m03;
        const double s3 = m01 * m12 - m11 * m02, s4 = m01 * m13 - m11 * m03, s5 = m02 * m13 - m12 * m03;
        const double c5 = m22 * m33 - m32 * m23, c4 = m21 * m33 - m31 * m23, c3 = m21 * m32 - m31 * m22;
        const double c2 = m20 * m33 - m30 * m23, c1 = m20 * m32 - m30 * m22, c0 = m20 * m31 - m30 * m21;
        const double invdet = 1.0 / (s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0);
        Mat4 r{};
        r.m[0][0] = (m11 * c5 - m12 * c4 + m13 * c3) * invdet;
        r.m[0][1] = (-m01 * c5 + m02 * c4 - m03 * c3) * invdet;
        r.m[0][2] = (m31 * s5 - m32 * s4 + m33 * s3) * invdet;
        r.m[0][3] = (-m21 * s5 + m22 * s4 - m23 * s3) * invdet;
        r.m[1][0] = (-m10 * c5 + m12 * c2 - m13 * c1) * invdet;
        r.m[1][1] = (m00 * c5 - m02 * c2 + m03 * c1) * invdet;
        r.m[1][2] = (-m30 * s5 + m32 * s2 - m33 * s1) * invdet;
        r.m[1][3] = (m20 * s5 - m22 * s2 + m23 * s1) * invdet;
        r.m[2][0] = (m10 * c4 - m11 * c2 + m13 * c0) * invdet;
        r.m[2][1] = (-m00 * c4 + m01 * c2 - m03 * c0) * invdet;
        r.m[2][2] = (m30 * s4 - m31 * s2 + m33 * s0) * invdet;
        r.m[2][3] = (-m20 * s4 + m21 * s2 - m23 * s0) * invdet;
        r.m[3][0] = (-m10 * c3 + m11 * c1 - m12 * c0) * invdet;
        r.m[3][1] = (m00 * c3 - m01 * c1 + m02 * c0) * invdet;
        r.m[3][2] = (-m30 * s3 + m31 * s1 - m32 * s0) * invdet;
        r.m[3][3] = (m20 * s3 - m21 * s1 + m22 * s0) * invdet;
        return r;
    }
};

// Vec3Ext, src/math.rs:36-52: Mat4 * (p, 1) and Mat4 * (d, 0), w dropped.
inline Vec3 transformed_point(Vec3 p, const Mat4& t) {
    return {t.m[0][0] * p.x + t.m[0][1] * p.y + t.m[0][2] * p.z + t.m[0][3],
            t.m[1][0] * p.x + t.m[1][1] * p.y + t.m[1][2] * p.z + t.m[1][3],
            t.m[2][0] * p.x + t.m[2][1] * p.y + t.m[2][2] * p.z + t.m[2][3]};
}
inline Vec3 transformed_direction(Vec3 d, const Mat4& t) {
    return {t.m[0][0] * d.x + t.m[0][1] * d.y + t.m[0][2] * d.z,
            t.m[1][0] * d.x + t.m[1][1] * d.y + t.m[1][2] * d.z,
            t.m[2][0] * d.x + t.m[2][1] * d.y + t.m[2][2] * d.z};
}

// Row-major 3x3 (vek Mat3::new takes arguments in row-major reading order, texture.rs:214-218)
struct Mat3 {
    double m[3][3];
    static Mat3 identity() {
        Mat3 r{};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[i][j] = (i == j) ? 1.0 : 0.0;
        return r;
    }
    static Mat3 scaling_3d(Vec3 v) {
        Mat3 r = identity();
        r.m[0][0] = v.x; r.m[1][1] = v.y; r.m[2][2] = v.z;
        return r;
    }
    Vec3 operator*(Vec3 v) const {
        return {m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z,
                m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z};
    }
};

}  // namespace portrayer
