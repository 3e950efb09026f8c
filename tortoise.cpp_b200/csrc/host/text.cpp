// text.cpp -- optional text front-end: the reference only accepts lower-case
// letters and `space . , ! ? ' -` (README.md:32; anything else falls through the greedy
// tokenizer as [UNK] or is silently dropped, common.cpp:320-336).  Both helpers are OFF by
// default everywhere (drop-in behaviour); the CLI enables normalisation with --normalize.
//   tts_host_normalize_text: ASCII lower-casing, cardinal numbers / decimals spelled out,
//       a few symbols (& % + = @ $) spelled out, everything else outside the supported
//       alphabet replaced by a space, runs of spaces collapsed.
//   tts_host_split_text: sentence-boundary chunking of long text into pieces of at most
//       max_chars characters (the AR stage is limited to 404 positions, main.cpp:794-797).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/tortoise_host.h"

namespace {

const char *kOnes[] = {"zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten",
                       "eleven", "twelve", "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen",
                       "nineteen"};
const char *kTens[] = {"", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"};

void below_thousand(unsigned n, std::string &out) {
  if (n >= 100) {
    out += kOnes[n / 100];
    out += " hundred";
    n %= 100;
    if (n) out += ' ';
  }
  if (n >= 20) {
    out += kTens[n / 10];
    if (n % 10) { out += '-'; out += kOnes[n % 10]; }
  } else if (n > 0) {
    out += kOnes[n];
  }
}

// digits (no sign, no separators) -> words; more than 15 digits are read digit by digit
std::string cardinal(const std::string &digits) {
  if (digits.size() > 15) {
    std::string out;
    for (char c : digits) { if (!out.empty()) out += ' '; out += kOnes[c - '0']; }
    return out;
  }
  unsigned long long v = 0;
  for (char c : digits) v = v * 10 + unsigned(c - '0');
  if (v == 0) return "zero";
  static const char *scale[] = {"", " thousand", " million", " billion", " trillion"};
  std::string parts[5];
  int n = 0;
  while (v > 0 && n < 5) {
    const unsigned g = unsigned(v % 1000);
    if (g) { below_thousand(g, parts[n]); parts[n] += scale[n]; }
    v /= 1000;
    ++n;
  }
  std::string out;
  for (int i = n - 1; i >= 0; --i)
    if (!parts[i].empty()) { if (!out.empty()) out += ' '; out += parts[i]; }
  return out;
}

bool supported(char c) {
  return (c >= 'a' && c <= 'z') || c == ' ' || c == '.' || c == ',' || c == '!' || c == '?' || c == '\'' || c == '-';
}

std::string normalize(const std::string &in) {
  std::string s;
  const size_t n = in.size();
  for (size_t i = 0; i < n;) {
    unsigned char c = static_cast<unsigned char>(in[i]);
    if (c >= '0' && c <= '9') {
      size_t j = i;
      while (j < n && in[j] >= '0' && in[j] <= '9') ++j;
      const bool dollars = i > 0 && in[i - 1] == '$';
      s += ' ';
      s += cardinal(in.substr(i, j - i));
      if (j + 1 < n && in[j] == '.' && in[j + 1] >= '0' && in[j + 1] <= '9') {  // decimal part, digit by digit
        s += " point";
        ++j;
        while (j < n && in[j] >= '0' && in[j] <= '9') { s += ' '; s += kOnes[in[j] - '0']; ++j; }
      }
      if (dollars) s += " dollars";
      s += ' ';
      i = j;
      continue;
    }
    if (c >= 'A' && c <= 'Z') c = static_cast<unsigned char>(c - 'A' + 'a');
    switch (c) {
      case '&': s += " and "; break;
      case '%': s += " percent "; break;
      case '+': s += " plus "; break;
      case '=': s += " equals "; break;
      case '@': s += " at "; break;
      default: s += supported(char(c)) ? char(c) : ' ';
    }
    ++i;
  }
  std::string out;
  for (char c : s) {  // collapse spaces, no space before punctuation
    if (c == ' ') {
      if (!out.empty() && out.back() != ' ') out += ' ';
    } else {
      if ((c == '.' || c == ',' || c == '!' || c == '?') && !out.empty() && out.back() == ' ') out.pop_back();
      out += c;
    }
  }
  while (!out.empty() && out.back() == ' ') out.pop_back();
  return out;
}

}  // namespace

extern "C" int tts_host_normalize_text(const char *in, char *out, int cap) {
  if (!in || !out || cap <= 0) return -1;
  const std::string r = normalize(in);
  if (int(r.size()) + 1 > cap) return -int(r.size()) - 1;  // -(bytes needed)
  memcpy(out, r.c_str(), r.size() + 1);
  return int(r.size());
}

extern "C" int tts_host_split_text(const char *in, int max_chars, int32_t *spans_out, int cap_spans) {
  if (!in || !spans_out || max_chars < 8 || cap_spans <= 0) return -1;
  const int n = int(strlen(in));
  int count = 0, pos = 0;
  while (pos < n) {
    while (pos < n && in[pos] == ' ') ++pos;
    if (pos >= n) break;
    int end = n;
    if (n - pos > max_chars) {
      const int limit = pos + max_chars;
      int cut = -1;
      for (int i = limit - 1; i > pos; --i)  // last sentence end inside the window
        if ((in[i] == '.' || in[i] == '!' || in[i] == '?') && (i + 1 >= n || in[i + 1] == ' ')) { cut = i + 1; break; }
      if (cut < 0)
        for (int i = limit - 1; i > pos; --i)  // else the last comma
          if (in[i] == ',') { cut = i + 1; break; }
      if (cut < 0)
        for (int i = limit; i > pos; --i)  // else the last space
          if (in[i] == ' ') { cut = i; break; }
      end = cut > pos ? cut : limit;  // a single over-long word is cut hard
    }
    int e = end;
    while (e > pos && in[e - 1] == ' ') --e;
    if (count >= cap_spans) return -2;
    spans_out[2 * count] = pos;
    spans_out[2 * count + 1] = e;
    ++count;
    pos = end;
  }
  return count;
}
