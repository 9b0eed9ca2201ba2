// Stand-in for <tf/tf.h> (TEST INFRASTRUCTURE, see oracle/refshim/README.md): tf::Quaternion and createQuaternionFromYaw
// as tf/LinearMath/Quaternion.h computes them (setRPY(0, 0, yaw) in double: x = y = 0, z = sin(yaw/2), w = cos(yaw/2)).
#ifndef REFSHIM_TF_H
#define REFSHIM_TF_H
#include <cmath>
namespace tf {
class Quaternion {
public:
    Quaternion() {}
    Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
    double x() const { return x_; }
    double y() const { return y_; }
    double z() const { return z_; }
    double w() const { return w_; }
private:
    double x_ = 0, y_ = 0, z_ = 0, w_ = 1;
};
inline Quaternion createQuaternionFromYaw(double yaw)
{
    const double halfYaw = yaw * 0.5, halfPitch = 0.0, halfRoll = 0.0;
    const double cosYaw = std::cos(halfYaw), sinYaw = std::sin(halfYaw), cosPitch = std::cos(halfPitch), sinPitch = std::sin(halfPitch),
                 cosRoll = std::cos(halfRoll), sinRoll = std::sin(halfRoll);
    return Quaternion(sinRoll * cosPitch * cosYaw - cosRoll * sinPitch * sinYaw, cosRoll * sinPitch * cosYaw + sinRoll * cosPitch * sinYaw,
                      cosRoll * cosPitch * sinYaw - sinRoll * sinPitch * cosYaw, cosRoll * cosPitch * cosYaw + sinRoll * sinPitch * sinYaw);
}
}  // namespace tf
#endif
