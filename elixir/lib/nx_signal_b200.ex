defmodule NxSignalB200.NIF do
  @moduledoc false
  # Source-only (no BEAM in this repository's environment).  Loads priv/nxsignal_nif.so built
  # from elixir/c_src/nxsignal_nif.c over include/nxsignal_b200.h.
  @on_load :load
  def load, do: :erlang.load_nif(~c"#{:code.priv_dir(:nx_signal_b200)}/nxsignal_nif", 0)
  def ctx_create(_dev), do: :erlang.nif_error(:not_loaded)
  def stft(_c, _x, _ch, _len, _w, _hop, _nfft, _pad, _lo, _hi, _scal, _sr), do: :erlang.nif_error(:not_loaded)
  def istft(_c, _z, _ch, _frames, _zlen, _w, _hop, _nfft, _scal, _sr), do: :erlang.nif_error(:not_loaded)
  def fir(_c, _x, _ch, _len, _h, _mode), do: :erlang.nif_error(:not_loaded)
  def stft_to_mel(_c, _z, _ch, _frames, _zlen, _nfft, _mels, _sr, _max_mel, _f_sp), do: :erlang.nif_error(:not_loaded)
  def stft_mel(_c, _x, _ch, _len, _w, _hop, _nfft, _pad, _lo, _hi, _scal, _sr, _mels, _max_mel, _f_sp),
    do: :erlang.nif_error(:not_loaded)

  def median(_c, _t, _shape, _kernel_shape), do: :erlang.nif_error(:not_loaded)
  def window(_kind, _n, _periodic, _beta, _eps), do: :erlang.nif_error(:not_loaded)
end

defmodule NxSignalB200 do
  @moduledoc """
  Drop-in heads for the accelerated path of `NxSignal`: `stft/3`, `istft/3` and the FIR form of
  `NxSignal.Convolution.convolve/3`, with the reference's options, defaults, error messages and
  return shapes (`lib/nx_signal.ex:68-130`, `:582-638`; `lib/nx_signal/convolution.ex:38-58`).

  Source-only: mirrors `nx_signal_b200/__init__.py`, which is the exercised host binding here.
  Inside a `defn` (tensors are `Nx.Defn.Expr`) the calls fall back to `NxSignal` itself.
  """

  @pad %{valid: 0, same: 1, reflect: 2}
  @scaling %{nil => 0, spectrum: 1, psd: 2}
  @mode %{full: 0, same: 1, valid: 2}

  defp ctx do
    case :persistent_term.get({__MODULE__, :ctx}, nil) do
      nil ->
        {:ok, c} = NxSignalB200.NIF.ctx_create(0)
        :persistent_term.put({__MODULE__, :ctx}, c)
        c

      c ->
        c
    end
  end

  defp expr?(%Nx.Tensor{data: %Nx.Defn.Expr{}}), do: true
  defp expr?(_), do: false

  defp raise_nif({:error, :argument_error, msg}), do: raise(ArgumentError, List.to_string(msg))
  defp raise_nif({:error, _, msg}), do: raise(RuntimeError, List.to_string(msg))

  def stft(data, window, opts \\ []) do
    if expr?(data) or expr?(window) do
      NxSignal.stft(data, window, opts)
    else
      {frame_length} = Nx.shape(window)

      opts =
        Keyword.validate!(opts, [
          :overlap_length,
          :window,
          :scaling,
          window_padding: :valid,
          sampling_rate: 100,
          fft_length: :power_of_two
        ])

      sampling_rate = opts[:sampling_rate] || raise ArgumentError, "missing sampling_rate option"
      overlap = opts[:overlap_length] || div(frame_length, 2)

      scaling =
        case Map.fetch(@scaling, opts[:scaling]) do
          {:ok, s} -> s
          :error -> raise ArgumentError, "invalid :scaling, expected one of :spectrum, :psd or nil, got: #{inspect(opts[:scaling])}"
        end

      {pad, lo, hi} =
        case opts[:window_padding] do
          p when is_map_key(@pad, p) -> {@pad[p], 0, 0}
          [{lo, hi}] when is_integer(lo) and is_integer(hi) -> {3, lo, hi}
          other -> raise ArgumentError, "invalid padding mode specified, padding must be one of :valid, :same, or a padding configuration, got: #{inspect(other)}"
        end

      nfft =
        case opts[:fft_length] do
          :power_of_two -> 2 ** ceil(:math.log2(frame_length))
          n -> n
        end

      vec_axes = data.vectorized_axes
      flat = data |> Nx.devectorize() |> Nx.as_type(:f32)
      len = Nx.axis_size(flat, -1)
      ch = div(Nx.size(flat), len)

      case NxSignalB200.NIF.stft(ctx(), Nx.to_binary(flat), ch, len, Nx.to_binary(Nx.as_type(window, :f32)),
             frame_length - overlap, nfft, pad, lo, hi, scaling, sampling_rate * 1.0) do
        {:ok, z, times, freqs, frames} ->
          z =
            z
            |> Nx.from_binary(:c64)
            |> Nx.reshape(Tuple.to_list(Nx.shape(flat)) |> List.replace_at(-1, frames) |> Kernel.++([nfft]) |> List.to_tuple())
            |> Nx.vectorize(vec_axes)
            |> then(&Nx.reshape(&1, &1.shape, names: [:frames, :frequencies]))

          {z, Nx.from_binary(times, :f32) |> Nx.reshape({frames}, names: [:frames]),
           Nx.from_binary(freqs, :f32) |> Nx.reshape({nfft}, names: [:frequencies])}

        err ->
          raise_nif(err)
      end
    end
  end

  def istft(data, window, opts) do
    if expr?(data) or expr?(window) do
      NxSignal.istft(data, window, opts)
    else
      opts = Keyword.validate!(opts, [:fft_length, :overlap_length, :scaling, sampling_rate: 1000])
      n = Nx.size(window)
      zlen = Nx.axis_size(data, -1)
      frames = Nx.axis_size(data, -2)
      nfft = opts[:fft_length] || 2 ** ceil(:math.log2(zlen))
      overlap = opts[:overlap_length] || div(n, 2)

      if opts[:scaling] == :psd and is_nil(opts[:sampling_rate]),
        do: raise(ArgumentError, ":sampling_rate is mandatory if scaling is :psd")

      scaling = Map.get(@scaling, opts[:scaling]) ||
        raise ArgumentError, "invalid :scaling, expected one of :spectrum, :psd or nil, got: #{inspect(opts[:scaling])}"

      vec_axes = data.vectorized_axes
      flat = data |> Nx.devectorize() |> Nx.as_type(:c64)
      ch = div(Nx.size(flat), frames * zlen)

      case NxSignalB200.NIF.istft(ctx(), Nx.to_binary(flat), ch, frames, zlen, Nx.to_binary(Nx.as_type(window, :f32)),
             n - overlap, nfft, scaling, (opts[:sampling_rate] || 1000) * 1.0) do
        {:ok, y} ->
          out_len = frames * (n - overlap) + overlap
          lead = flat |> Nx.shape() |> Tuple.to_list() |> Enum.drop(-2)
          y |> Nx.from_binary(:c64) |> Nx.reshape(List.to_tuple(lead ++ [out_len])) |> Nx.vectorize(vec_axes)

        err ->
          raise_nif(err)
      end
    end
  end

  @doc "FIR form of `NxSignal.Convolution.convolve/3`: `x {C, L}` with `taps {1, K}` (or both rank 1)."
  def fir(x, taps, opts \\ []) do
    opts = Keyword.validate!(opts, mode: :full, method: :direct)

    unless opts[:mode] in [:full, :same, :valid],
      do: raise(ArgumentError, "expected mode to be one of [:full, :same, :valid], got: #{inspect(opts[:mode])}")

    len = Nx.axis_size(x, -1)
    ch = div(Nx.size(x), len)

    case NxSignalB200.NIF.fir(ctx(), Nx.to_binary(Nx.as_type(x, :f32)), ch, len,
           Nx.to_binary(Nx.as_type(Nx.flatten(taps), :f32)), @mode[opts[:mode]]) do
      {:ok, y, out_len} ->
        shape = x |> Nx.shape() |> Tuple.to_list() |> List.replace_at(-1, out_len) |> List.to_tuple()
        y |> Nx.from_binary(:f32) |> Nx.reshape(shape)

      err ->
        raise_nif(err)
    end
  end

  @doc "`NxSignal.stft_to_mel/3` (lib/nx_signal.ex:486-513): `z {frames, frequencies}` (vectorised axes = batch)."
  def stft_to_mel(z, sampling_rate, opts \\ []) do
    opts = Keyword.validate!(opts, [:fft_length, :max_mel, :mel_frequency_spacing, mel_bins: 128, type: {:f, 32}])
    nfft = opts[:fft_length] || raise(ArgumentError, "missing :fft_length option")
    vec_axes = z.vectorized_axes
    zz = Nx.devectorize(z)
    frames = Nx.axis_size(zz, -2)
    zlen = Nx.axis_size(zz, -1)
    ch = div(Nx.size(zz), frames * zlen)

    case NxSignalB200.NIF.stft_to_mel(ctx(), Nx.to_binary(Nx.as_type(zz, :c64)), ch, frames, zlen, nfft,
           opts[:mel_bins], sampling_rate * 1.0, (opts[:max_mel] || 3016) * 1.0,
           (opts[:mel_frequency_spacing] || 200 / 3) * 1.0) do
      {:ok, mel} ->
        shape = zz |> Nx.shape() |> Tuple.to_list() |> Enum.drop(-2) |> Kernel.++([frames, opts[:mel_bins]]) |> List.to_tuple()
        mel |> Nx.from_binary(:f32) |> Nx.reshape(shape) |> Nx.vectorize(vec_axes) |> Nx.rename([:frames, :mel])

      err ->
        raise_nif(err)
    end
  end
  @doc """
  `NxSignal.stft/3 |> NxSignal.stft_to_mel/3` in one device call (the spectrum is never stored and only
  the `{frames, mel}` tensor crosses PCIe).  Options: those of `stft/3` plus `:mel_bins`, `:max_mel`,
  `:mel_frequency_spacing`; `:window_padding` is `:valid | :same | :reflect` here.
  """
  def stft_mel(data, window, opts \\ []) do
    {frame_length} = Nx.shape(window)

    opts =
      Keyword.validate!(opts, [
        :overlap_length,
        :scaling,
        :max_mel,
        :mel_frequency_spacing,
        window_padding: :valid,
        sampling_rate: 100,
        fft_length: :power_of_two,
        mel_bins: 128
      ])

    overlap = opts[:overlap_length] || div(frame_length, 2)
    nfft = if opts[:fft_length] == :power_of_two, do: next_pow2(frame_length), else: opts[:fft_length]
    vec_axes = data.vectorized_axes
    x = Nx.devectorize(data)
    len = Nx.axis_size(x, -1)
    ch = div(Nx.size(x), len)

    case NxSignalB200.NIF.stft_mel(ctx(), Nx.to_binary(Nx.as_type(x, :f32)), ch, len,
           Nx.to_binary(Nx.as_type(window, :f32)), frame_length - overlap, nfft, @pad[opts[:window_padding]], 0, 0,
           @scaling[opts[:scaling]], opts[:sampling_rate] * 1.0, opts[:mel_bins], (opts[:max_mel] || 3016) * 1.0,
           (opts[:mel_frequency_spacing] || 200 / 3) * 1.0) do
      {:ok, mel, frames} ->
        shape = x |> Nx.shape() |> Tuple.to_list() |> Enum.drop(-1) |> Kernel.++([frames, opts[:mel_bins]]) |> List.to_tuple()
        mel |> Nx.from_binary(:f32) |> Nx.reshape(shape) |> Nx.vectorize(vec_axes) |> Nx.rename([:frames, :mel])

      err ->
        raise_nif(err)
    end
  end

  defp next_pow2(n), do: Bitwise.bsl(1, ceil(:math.log2(n)))

  @doc "`NxSignal.Filters.median/2` (lib/nx_signal/filters.ex:17-56) for tensors of rank <= 3."
  def median(t, opts) do
    opts = Keyword.validate!(opts, [:kernel_shape])

    if Nx.rank(t) != tuple_size(opts[:kernel_shape]),
      do: raise(ArgumentError, "kernel shape must be of the same rank as the tensor")

    case NxSignalB200.NIF.median(ctx(), Nx.to_binary(Nx.as_type(t, :f32)), Tuple.to_list(Nx.shape(t)),
           Tuple.to_list(opts[:kernel_shape])) do
      {:ok, out} -> out |> Nx.from_binary(:f32) |> Nx.reshape(Nx.shape(t))
      err -> raise_nif(err)
    end
  end
end
