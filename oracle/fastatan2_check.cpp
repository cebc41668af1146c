// TEST INFRASTRUCTURE: prints the stand-in's cv::fastAtan2 (oracle/shim/opencv2/core.hpp) for "y x" pairs read from stdin, as float bit patterns,
// so that tests/test_hpatches_tool.py can pin it against cv2.fastAtan2.
#include <cstdio>
#include <cstring>
#include <opencv2/core.hpp>
int main()
{
    float y, x;
    while (std::scanf("%f %f", &y, &x) == 2) {
        const float a = cv::fastAtan2(y, x);
        unsigned u; std::memcpy(&u, &a, 4);
        std::printf("%u\n", u);
    }
    return 0;
}
