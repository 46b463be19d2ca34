//! Dumps what the reference computes for the committed inputs, as little-endian binaries:
//!   rust_in_<name>.bin     n interleaved (re, im) f64           (committed, written by tests/golden/make_golden.py)
//!   rust_fft_<name>.bin    kopek::fft::fft(input): next_power_of_two(n) interleaved (re, im) f64   (src/fft.rs:6-41)
//!   rust_show.txt          the "{:.4}{:+.4}i" bin format of kopek::fft::show (src/fft.rs:43-52) on tricky values
//!   rust_detect_bpm.txt    kopek::decoder::detect_bpm(decode(sample.wav), 44100.0)               (src/lib.rs:33-38)
//! usage: kopek-golden <tests/golden dir> [sample.wav]
use num::complex::Complex;
use std::fs;
use std::io::Write;
use std::path::Path;

fn read_complex(path: &Path) -> Vec<Complex<f64>> {
    let raw = fs::read(path).unwrap_or_else(|e| panic!("{}: {}", path.display(), e));
    assert!(raw.len() % 16 == 0, "{}: not a multiple of 16 bytes", path.display());
    raw.chunks_exact(16)
        .map(|c| Complex {
            re: f64::from_le_bytes(c[0..8].try_into().unwrap()),
            im: f64::from_le_bytes(c[8..16].try_into().unwrap()),
        })
        .collect()
}

fn write_complex(path: &Path, v: &[Complex<f64>]) {
    let mut f = fs::File::create(path).unwrap_or_else(|e| panic!("{}: {}", path.display(), e));
    for c in v {
        f.write_all(&c.re.to_le_bytes()).unwrap();
        f.write_all(&c.im.to_le_bytes()).unwrap();
    }
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let dir = Path::new(args.get(1).map(String::as_str).unwrap_or("tests/golden"));
    let mut done = 0;
    for entry in fs::read_dir(dir).expect("golden directory") {
        let path = entry.unwrap().path();
        let name = path.file_name().unwrap().to_string_lossy().into_owned();
        if let Some(stem) = name.strip_prefix("rust_in_").and_then(|s| s.strip_suffix(".bin")) {
            let input = read_complex(&path);
            let spectrum = kopek::fft::fft(&input);
            write_complex(&dir.join(format!("rust_fft_{}.bin", stem)), &spectrum);
            println!("{}: {} -> {} points", stem, input.len(), spectrum.len());
            done += 1;
        }
    }
    assert!(done > 0, "no rust_in_*.bin under {}", dir.display());

    // the bin format of show(): same expression as src/fft.rs:46-50
    let tricky = [
        Complex { re: 1.0, im: 2.0 },
        Complex { re: -0.5, im: -0.25 },
        Complex { re: 0.0, im: 0.0 },
        Complex { re: -0.0, im: -3.14159 },
        Complex { re: f64::NAN, im: f64::NEG_INFINITY },
        Complex { re: f64::INFINITY, im: 0.00004 },
        Complex { re: 123456.78905, im: -0.00005 },
    ];
    let line: Vec<String> = tricky.iter().map(|c| format!("{:.4}{:+.4}i", c.re, c.im)).collect();
    fs::write(dir.join("rust_show.txt"), format!("label\n{}\n", line.join(", "))).unwrap();

    if let Some(wav) = args.get(2) {
        let frames = kopek::decoder::decode(wav);
        let bpm = kopek::decoder::detect_bpm(frames, 44100.0);
        fs::write(dir.join("rust_detect_bpm.txt"), format!("{}\n", bpm)).unwrap();
        println!("detect_bpm({}) = {}", wav, bpm);
    }
}
