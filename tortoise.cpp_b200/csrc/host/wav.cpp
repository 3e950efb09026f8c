// wav.cpp -- 44-byte RIFF header + float32 mono samples, byte-identical to the reference's
// writeWav (main.cpp:4821-4868): format tag 3 (IEEE float), 32 bits, fileSize = 36 + bytes.
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace tts_host {
bool write_wav(const char *path, const float *data, int64_t n, int sample_rate) {
  FILE *f = fopen(path, "wb");
  if (!f) return false;
  const int32_t channels = 1, bits = 32;
  const int32_t byte_rate = sample_rate * channels * bits / 8, block_align = channels * bits / 8;
  const int32_t data_size = int32_t(n * int64_t(sizeof(float))), file_size = 36 + data_size, fmt_size = 16, tag = 3;
  fwrite("RIFF", 1, 4, f);
  fwrite(&file_size, 4, 1, f);
  fwrite("WAVE", 1, 4, f);
  fwrite("fmt ", 1, 4, f);
  fwrite(&fmt_size, 4, 1, f);
  fwrite(&tag, 2, 1, f);
  fwrite(&channels, 2, 1, f);
  fwrite(&sample_rate, 4, 1, f);
  fwrite(&byte_rate, 4, 1, f);
  fwrite(&block_align, 2, 1, f);
  fwrite(&bits, 2, 1, f);
  fwrite("data", 1, 4, f);
  fwrite(&data_size, 4, 1, f);
  fwrite(data, sizeof(float), size_t(n), f);
  fclose(f);
  return true;
}
}  // namespace tts_host
